#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mbconv.py -m gpu -q -x > gpurun_out/t_mb.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_mb.log; grep -E "^E  |FAILED|b200seg:" gpurun_out/t_mb.log | head -12 | cut -c1-300
timeout 600 python tools/kbench_mb.py 0 0x40 > gpurun_out/kb_mb.log 2>&1; echo "kbench rc=$?"; cat gpurun_out/kb_mb.log | tail -14
