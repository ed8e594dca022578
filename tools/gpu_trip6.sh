#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
python tools/ncu_cases.py tc,dw,upcat,final,stem > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"dwconv|upsample|conv3x3_c3|conv_tc" -c 30 -o gpurun_out/prof_r1b python tools/ncu_cases.py tc,dw,upcat,final,stem > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
cat gpurun_out/bench.json
