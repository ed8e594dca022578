#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_dp_gpu.py -m gpu -q -x > gpurun_out/t_train.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_train.log; grep -E "^E  |FAILED|Error" gpurun_out/t_train.log | head -8 | cut -c1-300
timeout 600 python bench.py --workload train --steps 10 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['host_issue_ms_per_step'])"
timeout 900 python bench.py --workload unet_train --steps 4 --warmup 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('unet', d['ms_per_step'], d['value'])"
timeout 600 python tools/train_profile.py 2>&1 | grep -E "total device"
