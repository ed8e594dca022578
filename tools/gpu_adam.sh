#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "adam or graph" > gpurun_out/t_adam.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_adam.log; grep -E "^E  |FAILED" gpurun_out/t_adam.log | head -8 | cut -c1-300
for o in torch fused; do
B200SEG_BENCH_ADAM=$o timeout 600 python bench.py --workload train --steps 10 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$o', d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['host_issue_ms_per_step'])"
done
