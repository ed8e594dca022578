#!/bin/bash
# Round-end validation: full GPU test suite, smoke, both bench arms, training bench, launch list, ncu of the dominant kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_notes.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_all.log; grep -E "^E  |FAILED" gpurun_out/t_all.log | head -8 | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 600 python bench.py --workload train --steps 10 --warmup 5 --breakdown > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train bench rc=$?"; cut -c1-200 gpurun_out/bench_train.json
timeout 600 python bench.py --workload infer720 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/bench_infer720.json 2> gpurun_out/bench_infer720.err; echo "infer720 rc=$?"; cut -c1-200 gpurun_out/bench_infer720.json
timeout 900 python bench.py --workload unet_train --steps 4 --warmup 4 > gpurun_out/bench_unet_train.json 2> gpurun_out/bench_unet_train.err; echo "unet_train rc=$?"; cut -c1-200 gpurun_out/bench_unet_train.json
timeout 600 python tools/train_profile.py > gpurun_out/train_profile.txt 2>&1; echo "profile rc=$?"
