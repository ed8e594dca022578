#!/bin/bash
timeout 300 python tools/host_profile_train.py 2>&1 | tail -60
