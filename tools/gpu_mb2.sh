#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_mbconv.py -m gpu -q -x > gpurun_out/t_mb.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t_mb.log; grep -E "^E  |FAILED|b200seg:" gpurun_out/t_mb.log | head -12 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --breakdown > gpurun_out/bench_mb.json 2> gpurun_out/bench_mb.err; echo "bench rc=$?"; cut -c1-250 gpurun_out/bench_mb.json; head -30 gpurun_out/bench_mb.err | cut -c1-110; tail -1 gpurun_out/bench_mb.err
