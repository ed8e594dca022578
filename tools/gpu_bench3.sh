#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8 9 10; do
B200SEG_DEBUG_ISSUE=1 timeout 900 python bench.py --no-cpu-baseline --sustain-seconds 0.1 > gpurun_out/r2_bench_$i.json 2> gpurun_out/r2_bench_$i.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_$i.json"))
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]))
PY
grep "dbg" gpurun_out/r2_bench_$i.err | tail -4 | cut -c1-200
done
