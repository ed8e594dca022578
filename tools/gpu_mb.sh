#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mbconv.py -m gpu -q -x 2>&1 | tail -8
timeout 300 python tools/kbench_mb.py 0 2>&1 | tail -14
