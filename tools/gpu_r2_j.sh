#!/bin/bash
# per-layer table of the training step (warm eager pass) + the inference table
mkdir -p gpurun_out
timeout 600 python bench.py --workload train --steps 10 --warmup 5 --breakdown > gpurun_out/r2_train_bd.json 2> gpurun_out/r2_train_bd.txt; echo "rc=$?"
head -c 400 gpurun_out/r2_train_bd.json
