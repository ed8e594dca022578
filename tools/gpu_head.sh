#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "stem_mb1" 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_f2_fixture.py -m gpu -q -x 2>&1 | tail -8
timeout 900 python bench.py --breakdown --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench.json"))
c=d["config"]
print("value",d["value"],"e2e",d["e2e"]["value"],"ms",d["ms_per_step"],"train",c.get("train_img_s"),c.get("train_ms_per_step"),"roof",d["roofline"]["kernel"],d["roofline"]["frac"])
PY
grep -A12 "^kernel" gpurun_out/r2_bench.err | cut -c1-100
