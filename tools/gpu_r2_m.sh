#!/bin/bash
# all GPU tests + the default bench line without the CPU leg
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/t_all.log; grep -E "^E  |FAILED" gpurun_out/t_all.log | head -8 | cut -c1-300
timeout 900 python bench.py --breakdown --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench.json")); c=d["config"]
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],4),"roof",d["roofline"]["kernel"],round(d["roofline"]["frac"],3), "train", round(c.get("train_img_s",0)), round(c.get("train_ms_per_step",0),3))
PY
grep -E "upcat|^sum" gpurun_out/r2_bench.err | cut -c1-110
