#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/ncu_model.py > gpurun_out/ncu_model_plain.log 2>&1 || exit 1
tail -1 gpurun_out/ncu_model_plain.log | cut -c1-200
# second forward only: 1 fold/pack launch + 31 kernels of the first forward are skipped
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:conv_rs|conv_tc|fold_pack|mbconv|stem_mb1|tail_fused|upsample2x" --launch-skip 32 --launch-count 31 -f -o gpurun_out/r02_forward python tools/ncu_model.py > gpurun_out/ncu_model.log 2>&1
tail -2 gpurun_out/ncu_model.log
