#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/ncu_bn_mid.py || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bn_ --launch-skip 12 --launch-count 6 -f -o gpurun_out/r02_bn python tools/ncu_bn_mid.py > gpurun_out/ncu_bn.log 2>&1
tail -2 gpurun_out/ncu_bn.log
