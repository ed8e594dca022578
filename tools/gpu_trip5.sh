#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in 13 22 23 24 43; do
  B200SEG_DW_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "test_dwconv3x3 and not tensor_core" 2>&1 | tail -1
  echo "== DW_VARIANT=$v"; B200SEG_DW_VARIANT=$v timeout 300 python tools/kbench.py dwconv 2>&1 | grep -v TENSOR
done > gpurun_out/dw_variants.log 2>&1
