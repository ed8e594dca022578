#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x > gpurun_out/t_fwd.log 2>&1; echo "fwd tests rc=$?"; tail -2 gpurun_out/t_fwd.log
timeout 300 python tools/kbench.py stem 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --breakdown --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
head -16 gpurun_out/bench.err; tail -1 gpurun_out/bench.err; cut -c1-220 gpurun_out/bench.json
