#!/bin/bash
# training-path regression + step time + kernel table after a kernel change
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_f2_fixture.py tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -6
timeout 600 python bench.py --workload train --steps 20 --warmup 5 > gpurun_out/r2_train.json 2> gpurun_out/r2_train.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_train.json"))
print("train value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"], d["config"].get("host_issue_ms_per_step"), d.get("launches_per_step"))
PY
TP_DETAIL=${TP_DETAIL:-bn_bwd_reduce,bn_bwd_apply,bn_stats,bn_apply_kernel} timeout 300 python tools/train_profile.py > gpurun_out/r2_train_profile.txt 2>&1; grep -v Warn gpurun_out/r2_train_profile.txt | head -24 | cut -c1-120
