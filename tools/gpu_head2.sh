#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "stem_mb1" 2>&1 | tail -3
timeout 200 python tools/ncu_head.py
