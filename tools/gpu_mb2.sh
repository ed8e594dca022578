#!/bin/bash
timeout 300 python tools/kbench_mb.py 0 0x20000 0x10000 2>&1 | tail -13 | cut -c1-250
