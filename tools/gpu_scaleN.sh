#!/bin/bash
# N-GPU run of the default bench line (inference value/e2e + training leg in config.train_*), launched exactly as the driver does
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_${N}gpu.json").read().strip().splitlines()[-1]); c=d["config"]
print("N",d["n_gpus"],"value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3),"train",round(c["train_img_s"]),round(c["train_ms_per_step"],2),"exposed",c.get("allreduce_exposed_ms"),"host",round(c.get("train_host_issue_ms"),2))
PY
grep -c "AllReduce" gpurun_out/r02_bench_${N}gpu.err; grep -m3 "NVLS\|Using network\|Channel" gpurun_out/r02_bench_${N}gpu.err | cut -c1-160
