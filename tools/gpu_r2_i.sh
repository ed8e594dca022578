#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/ncu_mb.py > gpurun_out/ncu_mb_plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mbconv -c 4 -f -o gpurun_out/r02_mb python tools/ncu_mb.py > gpurun_out/ncu_mb.log 2>&1
tail -3 gpurun_out/ncu_mb.log
