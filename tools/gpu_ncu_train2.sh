#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/ncu_train2.py > gpurun_out/nt2_plain.log 2>&1 || { tail -5 gpurun_out/nt2_plain.log; exit 1; }
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"upcat_bwd|dw_wgrad|dwconv3x3_bf16w|bn_stats|bn_apply|bn_bwd|smallcin3" -f -o gpurun_out/r02_train2 python tools/ncu_train2.py > gpurun_out/nt2.log 2>&1
tail -2 gpurun_out/nt2.log
ncu -i gpurun_out/r02_train2.ncu-rep --page raw --csv > gpurun_out/r02_train2_raw.csv 2>/dev/null; wc -l gpurun_out/r02_train2_raw.csv
