"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv` of bench.py): finds the inference value-leg steps
(31 launches from stem_mb1 to tail_fused) and prints per-kernel shares of one step plus its top single launches."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
names = [re.sub(r"^void ", "", r[4]).split("(")[0] for r in rows]
us = [float(r[14].replace(",", "")) / 1e3 for r in rows]
starts = [i for i, n in enumerate(names) if "stem_mb1_kernel" in n]
steps = [(a, a + 31) for a in starts if a + 31 <= len(names) and "tail_fused" in names[a + 30]]
print(f"launches captured: {len(rows)}; inference steps (stem_mb1 .. tail_fused, 31 launches) found: {len(steps)}")
use = steps[-4:]
agg = collections.OrderedDict()
for a, b in use:
    for i in range(a, b):
        k = names[i][:64]
        agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += us[i]
tot = sum(v[1] for v in agg.values())
print(f"\nlast {len(use)} inference steps of the capture (per-launch times are cold-cache and serialised by ncu: compare SHARES, not absolutes)")
print(f"{'kernel':66s} {'launches':>8s} {'total_us':>10s} {'share':>7s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:66s} {n:8d} {t:10.1f} {t / tot:7.1%}")
print(f"{'sum':66s} {sum(v[0] for v in agg.values()):8d} {tot:10.1f}")
a, b = use[-1]
print("\nTop 10 single launches of one step:")
for i in sorted(range(a, b), key=lambda i: -us[i])[:10]:
    print(f"  step-launch #{i - a:2d} {names[i][:60]:60s} {us[i]:8.1f} us {us[i] / sum(us[a:b]):6.1%}")
tr = [i for i, n in enumerate(names) if "softmax_ce" in n]
print(f"\ntraining-leg launches in the capture: {sum(1 for n in names if 'bn_bwd' in n)} BatchNorm-backward, {len(tr)} softmax_ce")
