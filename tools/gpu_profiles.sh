#!/bin/bash
# refresh the round's evidence files (copied from gpurun_out/ into profiles/ by hand)
mkdir -p gpurun_out
timeout 900 python bench.py --breakdown > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench_breakdown.txt || exit 1
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench.json")); c=d["config"]
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3),"train",round(c["train_img_s"]),round(c["train_ms_per_step"],2),"roof",d["roofline"]["kernel"],round(d["roofline"]["frac"],3), "step frac", round(d["step_roofline"]["frac"],3), "cpu", d["cpu_baseline"]["value"])
PY
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; tail -c 600 gpurun_out/r02_bench_reference_arm.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_ll.log 2>&1
wc -l gpurun_out/r02_launches.csv
