#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py tests/test_f2_fixture.py -m gpu -q -x 2>&1 | tail -4
timeout 900 python bench.py --breakdown --no-cpu-baseline --no-train-leg > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench.json"))
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],4),"roof",d["roofline"]["kernel"],round(d["roofline"]["frac"],3))
PY
grep -E "upcat|tail|^sum" gpurun_out/r2_bench.err | cut -c1-110
