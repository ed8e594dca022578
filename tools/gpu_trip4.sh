#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/t_all.log
for v in 0 1 2 3 9; do echo "== DW_VARIANT=$v"; B200SEG_DW_VARIANT=$v timeout 300 python tools/kbench.py dwconv 2>&1 | grep -v TENSOR; done > gpurun_out/dw_variants.log 2>&1
timeout 300 python tools/kbench.py upcat > gpurun_out/kb_upcat.log 2>&1
timeout 600 python tools/kbench.py conv_tc > gpurun_out/kbench.log 2>&1; echo "kbench rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json | cut -c1-300
