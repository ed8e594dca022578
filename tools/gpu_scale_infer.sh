#!/bin/bash
# inference bench under torch.distributed.run on N GPUs (value and e2e) for both e2e download variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
for v in 1 0 1 0; do
B200SEG_E2E_D2H_STREAM=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$v bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/scale_infer_$N.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N d2h_stream=$v', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
nproc; numactl -H 2>/dev/null | head -4; nvidia-smi topo -m 2>/dev/null | head -12
