#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -k "wgrad_tensor_core" > gpurun_out/t_wg.log 2>&1; echo "wgrad tc tests rc=$?"; grep -E "^E  .*Assertion|passed|failed" gpurun_out/t_wg.log | cut -c1-400 | head -20
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q -k "not wgrad_tensor_core" > gpurun_out/t_train.log 2>&1; echo "train tests rc=$?"; tail -3 gpurun_out/t_train.log
timeout 600 python bench.py --workload train --steps 3 --warmup 2 --breakdown > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train bench rc=$?"
grep -E "^(bwd|fwd)" gpurun_out/bench_train.err | head -12; tail -1 gpurun_out/bench_train.err; cat gpurun_out/bench_train.json | cut -c1-250
