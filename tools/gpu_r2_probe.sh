#!/bin/bash
# round-2 first call: instruction-rate probes + baseline numbers on today's box
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2_box.txt 2>&1
timeout 60 tools/_build/fhfma_probe > gpurun_out/r2_fhfma_probe.txt 2>&1
timeout 300 python tools/mma_probe.py > gpurun_out/r2_mma_probe.txt 2>&1
timeout 300 python bench.py --breakdown > gpurun_out/r2_base_infer.json 2> gpurun_out/r2_base_infer.err
timeout 300 python bench.py --workload train --breakdown > gpurun_out/r2_base_train.json 2> gpurun_out/r2_base_train.err
tail -3 gpurun_out/r2_fhfma_probe.txt; cat gpurun_out/r2_base_infer.json | cut -c1-300
