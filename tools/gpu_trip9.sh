#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --workload train --steps 3 --warmup 2 --breakdown > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train bench rc=$?"
tail -5 gpurun_out/bench_train.err; cat gpurun_out/bench_train.json | cut -c1-900
