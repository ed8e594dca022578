#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -120 > gpurun_out/r2_pytest.txt
tail -15 gpurun_out/r2_pytest.txt
timeout 300 python bench.py --workload train --breakdown > gpurun_out/r2_train.json 2> gpurun_out/r2_train.err
cut -c1-400 gpurun_out/r2_train.json; tail -3 gpurun_out/r2_train.err
