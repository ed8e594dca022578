#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_pytest.txt
tail -6 gpurun_out/r2_pytest.txt
timeout 600 python bench.py --breakdown --no-train-leg --no-eager-baseline --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"])
PY
head -14 gpurun_out/r2_bench.err; grep "tail\|final\|outc" gpurun_out/r2_bench.err
