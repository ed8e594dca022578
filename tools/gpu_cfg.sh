#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --workload infer720 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/bench_infer720.json 2> gpurun_out/bench_infer720.err; echo "infer720 rc=$?"; cut -c1-260 gpurun_out/bench_infer720.json; tail -2 gpurun_out/bench_infer720.err
timeout 900 python bench.py --workload unet_train --steps 4 --warmup 3 --breakdown > gpurun_out/bench_unet_train.json 2> gpurun_out/bench_unet_train.err; echo "unet_train rc=$?"; cut -c1-260 gpurun_out/bench_unet_train.json; tail -3 gpurun_out/bench_unet_train.err
