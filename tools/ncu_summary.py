"""Summarise an `ncu --page raw --csv` dump (argv[1]) of tools/ncu_model.py: one line per launch with the schedule step name
(argv[2] = file holding the SCHEDULE line), time, DRAM bytes, DRAM / tensor / issue utilisation, occupancy, registers."""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
sched = []
for line in open(sys.argv[2]):
    if line.startswith("SCHEDULE "):
        sched = line[9:].strip().split("|")
def g(r, k):
    try:
        return float(r[ix[k]].replace(",", ""))
    except (KeyError, ValueError):
        return float("nan")
unit_mb = lambda r, k: g(r, k) * {"Mbyte": 1.0, "Kbyte": 1e-3, "Gbyte": 1e3, "byte": 1e-6}.get(rows[1][ix[k]], 1.0)
out = {}
print(f"{'step':34s} {'kernel':26s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'dram%':>6s} {'tens%':>6s} {'issue%':>6s} {'occ%':>5s} {'regs':>4s}")
data = data[-len(sched):] if sched else data
for i, r in enumerate(data):
    name = r[ix["Kernel Name"]].split("(")[0].replace("void b200::", "").replace("<unnamed>::", "")[:26]
    step = sched[i] if i < len(sched) else "?"
    rd, wr = unit_mb(r, "dram__bytes_read.sum"), unit_mb(r, "dram__bytes_write.sum")
    print(f"{step[:34]:34s} {name:26s} {g(r, 'gpu__time_duration.sum'):8.1f} {rd:8.1f} {wr:8.1f} "
          f"{g(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'):6.1f} {g(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{g(r, 'sm__inst_issued.avg.pct_of_peak_sustained_active'):6.1f} {g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f} "
          f"{int(g(r, 'launch__registers_per_thread')):4d}")
    out[step.split(":", 1)[-1]] = {"traffic_bytes": int((rd + wr) * 1e6), "dram_read_bytes": int(rd * 1e6), "dram_write_bytes": int(wr * 1e6),
                                   "ncu_time_us": g(r, "gpu__time_duration.sum")}
if len(sys.argv) > 3:
    json.dump(out, open(sys.argv[3], "w"), indent=1)
