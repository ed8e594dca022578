#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_ops.py -m gpu -q -x -k "dw or depthwise or wgrad" 2>&1 | tail -3
for nm in 0 48; do for bps in 1 2; do echo "== NARROW_MB=$nm DWW_BPS=$bps"; B200SEG_DWW_NARROW_MB=$nm B200SEG_DWW_BPS=$bps timeout 120 python tools/kbench_train.py dw 2>&1 | grep dw_wgrad; done; done
