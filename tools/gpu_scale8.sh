#!/bin/bash
# 8-GPU weak-scaling check of both benches (launched exactly as the driver does)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${N}gpu_infer.json 2> gpurun_out/bench_${N}gpu_infer.err; echo "infer rc=$?"; cut -c1-220 gpurun_out/bench_${N}gpu_infer.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --workload train --steps 10 --warmup 5 > gpurun_out/bench_${N}gpu_train.json 2> gpurun_out/bench_${N}gpu_train.err; echo "train rc=$?"; cut -c1-220 gpurun_out/bench_${N}gpu_train.json; tail -3 gpurun_out/bench_${N}gpu_train.err | cut -c1-200
