#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/ncu_head.py > gpurun_out/ncu_head_plain.log 2>&1 || exit 1
cat gpurun_out/ncu_head_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stem_mb1 -s 2 -c 1 -f -o gpurun_out/r02_head python tools/ncu_head.py > gpurun_out/ncu_head.log 2>&1
tail -2 gpurun_out/ncu_head.log
