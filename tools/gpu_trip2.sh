#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for fl in 0 16 2 4 1; do
  timeout 300 python tools/tc_probe.py $fl > gpurun_out/probe_$fl.log 2>&1; echo "probe flags=$fl rc=$? bad=$(grep -c BAD gpurun_out/probe_$fl.log) ok=$(grep -c ' OK' gpurun_out/probe_$fl.log)"
done
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t_all.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python tools/kbench.py > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json | cut -c1-400
