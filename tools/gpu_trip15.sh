#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x > gpurun_out/t_train.log 2>&1; echo "train tests rc=$?"; tail -2 gpurun_out/t_train.log; grep -E "^E  " gpurun_out/t_train.log | head -5 | cut -c1-300
timeout 600 python tools/train_profile.py > gpurun_out/train_profile.txt 2>&1; head -16 gpurun_out/train_profile.txt | tail -14 | cut -c1-110
timeout 600 python bench.py --workload train --steps 8 --warmup 4 > gpurun_out/bench_train.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/bench_train.json')); print('ms/step', d['ms_per_step'], 'host issue ms', d['config']['host_issue_ms_per_step'], 'img/s', d['value'], 'e2e', d['e2e']['value'])"
