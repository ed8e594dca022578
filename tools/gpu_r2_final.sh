#!/bin/bash
# Round-2 evidence refresh: full GPU suite, smoke, both bench arms, training bench + tables, ncu launch list of the bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_notes.jsonl
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_all.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/t_all.log; grep -E "^E  |FAILED" gpurun_out/t_all.log | head -8 | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --breakdown > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench_breakdown.txt; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench.json")); c=d["config"]
print("value",round(d["value"]),"e2e",round(d["e2e"]["value"]),"ms",round(d["ms_per_step"],3),"train",round(c["train_img_s"]),round(c["train_ms_per_step"],2),"roof",d["roofline"]["kernel"],round(d["roofline"]["frac"],3), "step frac", round(d["step_roofline"]["frac"],3), "cpu", d["cpu_baseline"]["value"])
PY
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref rc=$?"; tail -c 300 gpurun_out/r02_bench_reference_arm.json
timeout 600 python bench.py --workload train --steps 20 --warmup 5 --breakdown > gpurun_out/r02_bench_config2_train.json 2> gpurun_out/r02_bench_config2_train_breakdown.txt; echo "train rc=$?"; cut -c1-160 gpurun_out/r02_bench_config2_train.json
timeout 300 python tools/train_profile.py 2>&1 | grep -v -i warn > gpurun_out/r02_train_step_kernel_profile.txt; head -3 gpurun_out/r02_train_step_kernel_profile.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-train-leg --no-eager-baseline > gpurun_out/ncu_ll.log 2>&1; echo "ncu rc=$?"; wc -l gpurun_out/r02_launches.csv
