#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "conv_tc or tensor_core" > gpurun_out/t_tc.log 2>&1; echo "tc tests rc=$?"; tail -3 gpurun_out/t_tc.log; grep -E "^E  " gpurun_out/t_tc.log | head -5 | cut -c1-300
timeout 300 python tools/kbench.py "k9" 2>&1 | sed -E 's/ +/ /g' | cut -d' ' -f2-5,7-13
