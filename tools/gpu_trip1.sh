#!/bin/bash
# first GPU trip: probe, tests (separate processes so a trapped kernel cannot poison later stages), smoke, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python tools/tc_probe.py > gpurun_out/probe_tma.log 2>&1; echo "probe tma rc=$?"
timeout 300 python tools/tc_probe.py direct > gpurun_out/probe_direct.log 2>&1; echo "probe direct rc=$?"
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "not conv_tc" > gpurun_out/t_ops.log 2>&1; echo "ops rc=$?"
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "conv_tc" > gpurun_out/t_tc.log 2>&1; echo "tc rc=$?"
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q > gpurun_out/t_model.log 2>&1; echo "model rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -3 gpurun_out/probe_tma.log; tail -3 gpurun_out/t_ops.log; tail -3 gpurun_out/t_tc.log; tail -3 gpurun_out/t_model.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json
