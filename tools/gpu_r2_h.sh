#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_pytest.txt
tail -6 gpurun_out/r2_pytest.txt
timeout 900 python bench.py --breakdown > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench.json"))
c=d["config"]
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],"train",c.get("train_img_s"),c.get("train_ms_per_step"),"roof",d["roofline"]["kernel"],d["roofline"]["frac"])
PY
grep -A12 "^kernel" gpurun_out/r2_bench.err | cut -c1-100
