#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_pytest.txt
tail -5 gpurun_out/r2_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --breakdown > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench.json"))
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"])
print({k:v for k,v in d["config"].items() if k.startswith(("train","eager","allreduce","sustained"))})
print(d.get("cpu_baseline"))
PY
tail -4 gpurun_out/r2_bench.err
