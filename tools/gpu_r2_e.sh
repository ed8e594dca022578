#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_pytest.txt
tail -6 gpurun_out/r2_pytest.txt
(cd tools && timeout 300 python kbench_mb.py 2>&1 | tail -14) | cut -c1-110
timeout 300 python tools/ab_eval.py dw_impl=simt dw_impl=None 2>&1 | tail -2
timeout 600 python bench.py --breakdown --no-eager-baseline --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"])
print({k:v for k,v in d["config"].items() if k.startswith(("train","sustained"))})
PY
