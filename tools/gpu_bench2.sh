#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2_bench_$i.json 2> gpurun_out/r2_bench_$i.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_$i.json"))
c=d["config"]
print("value",round(d["value"]),"sus",round(c["sustained_img_s"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3),"train",round(c.get("train_img_s")),round(c.get("train_ms_per_step"),2))
PY
done
