#!/bin/bash
# inference bench under torch.distributed.run on N GPUs (value and e2e), launched as the driver does
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/scale_infer_$N.err > gpurun_out/bench_${N}gpu_infer.json
python -c "import json; d=json.load(open('gpurun_out/bench_${N}gpu_infer.json')); print('N=$N', d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])"
