#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_notes.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_all.log; grep -E "^E  |FAILED" gpurun_out/t_all.log | head -8 | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 600 python bench.py --workload train --steps 10 --warmup 4 --breakdown > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train bench rc=$?"; cut -c1-200 gpurun_out/bench_train.json
python tools/ncu_dominant.py > gpurun_out/nd_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|dwconv3x3_bf16" -s 2 -c 1 -o gpurun_out/prof_dominant python tools/ncu_dominant.py > gpurun_out/nd.log 2>&1; echo "ncu rc=$?"
