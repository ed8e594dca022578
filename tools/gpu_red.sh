#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_ops.py -m gpu -q -x > gpurun_out/t_train.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_train.log; grep -E "^E  |FAILED" gpurun_out/t_train.log | head -8 | cut -c1-300
timeout 600 python bench.py --workload train --steps 10 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"
timeout 600 python tools/train_profile.py > gpurun_out/train_profile.txt 2>&1; grep -v Warn gpurun_out/train_profile.txt | head -34 | cut -c1-130
