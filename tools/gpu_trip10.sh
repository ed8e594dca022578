#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=29512
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "infer n2 rc=$?"; cat gpurun_out/bench_n2.json | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --workload train --steps 3 --warmup 2 > gpurun_out/bench_train_n2.json 2> gpurun_out/bench_train_n2.err; echo "train n2 rc=$?"; cat gpurun_out/bench_train_n2.json | cut -c1-200; tail -3 gpurun_out/bench_train_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --gpus 2 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n2.json 2>/dev/null; echo "ref n2 rc=$?"; cat gpurun_out/bench_ref_n2.json | cut -c1-150
timeout 600 python tools/dp_check.py > gpurun_out/dp_check.log 2>&1; echo "dp_check rc=$?"; tail -3 gpurun_out/dp_check.log
