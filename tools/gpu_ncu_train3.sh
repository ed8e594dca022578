#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/ncu_train3.py > gpurun_out/nt3_plain.log 2>&1 || { tail -5 gpurun_out/nt3_plain.log; exit 1; }
timeout 150 ncu --set full --clock-control none -k regex:"softmax_ce_x4|stem_wgrad|bn_stats|bn_apply|bn_bwd|conv_wgrad_tc|conv_rs" --launch-skip 7 --launch-count 7 -f -o gpurun_out/r02_train3 python tools/ncu_train3.py > gpurun_out/nt3.log 2>&1
tail -2 gpurun_out/nt3.log
ncu -i gpurun_out/r02_train3.ncu-rep --page raw --csv > gpurun_out/r02_train3_raw.csv 2>/dev/null; wc -l gpurun_out/r02_train3_raw.csv
