#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_train.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/train_profile.py 2>&1 | tail -45 > gpurun_out/r2_train_profile.txt; head -12 gpurun_out/r2_train_profile.txt | cut -c1-110
timeout 600 python bench.py --workload train > gpurun_out/r2_train.json 2> gpurun_out/r2_train.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_train.json"))
print("train value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"], d["config"]["host_issue_ms_per_step"], d["launches_per_step"])
PY
