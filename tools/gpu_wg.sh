#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x > gpurun_out/t_train.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_train.log; grep -E "^E  |FAILED|b200seg:" gpurun_out/t_train.log | head -8 | cut -c1-300
for r in 0 1; do
B200SEG_WGRAD_ROW=$r timeout 600 python bench.py --workload train --steps 10 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ROW=$r', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
B200SEG_WGRAD_ROW=1 timeout 900 python bench.py --workload unet_train --steps 4 --warmup 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('unet', d['ms_per_step'], d['value'])"
