#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/ncu_bn_mid.py || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bn_ --launch-skip 4 --launch-count 2 -f -o gpurun_out/r02_bnc python tools/ncu_bn_mid.py > gpurun_out/ncu_bnc.log 2>&1
tail -2 gpurun_out/ncu_bnc.log
ncu -i gpurun_out/r02_bnc.ncu-rep --page raw --csv > gpurun_out/r02_bnc_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_bnc.ncu-rep --page source --csv > gpurun_out/r02_bnc_src.csv 2>/dev/null; wc -l gpurun_out/r02_bnc_raw.csv gpurun_out/r02_bnc_src.csv
