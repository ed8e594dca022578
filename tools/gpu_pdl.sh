#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for p in 0 1 2 3 0 3; do
  B200SEG_PDL=$p timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pdl$p.json 2> gpurun_out/bench_pdl$p.err; echo "pdl=$p rc=$? $(python -c "import json;d=json.load(open('gpurun_out/bench_pdl$p.json'));print(d['ms_per_step'], d['e2e']['value'])")"
done
B200SEG_PDL=0 timeout 600 python tools/train_profile.py > gpurun_out/train_profile.txt 2>&1; echo "profile rc=$?"
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x > gpurun_out/t_train.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_train.log
