#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_pytest.txt
tail -6 gpurun_out/r2_pytest.txt
(cd tools && timeout 200 python kbench_train.py dw 2>&1 | grep dw_wgrad; B200SEG_DW_WGRAD=4 timeout 200 python kbench_train.py dw 2>&1 | grep dw_wgrad; B200SEG_DW_WGRAD=0 timeout 200 python kbench_train.py dw 2>&1 | grep dw_wgrad)
timeout 600 python bench.py --workload train --breakdown > gpurun_out/r2_train.json 2> gpurun_out/r2_train.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_train.json"))
print("train value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"], d["config"]["host_issue_ms_per_step"], d["launches_per_step"])
PY
tail -3 gpurun_out/r2_train.err
timeout 300 python tools/train_profile.py 2>&1 | tail -45 > gpurun_out/r2_train_profile.txt; head -30 gpurun_out/r2_train_profile.txt
