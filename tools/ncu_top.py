"""Top stalled SASS instructions of each kernel in an `ncu --page source --csv` dump (argv[1]); argv[2] = how many."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
only = [int(a) for a in sys.argv[3:]]
kern, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'rows': []}; kern.append(cur); continue
    if r and r[0] == 'Address':
        cur['hdr'] = r; continue
    if cur is not None and len(r) > 5:
        cur['rows'].append(r)
ST = ('stall_barrier', 'stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_mio', 'stall_not_selected', 'stall_selected',
      'stall_lg', 'stall_branch_resolving', 'stall_dispatch', 'stall_membar', 'stall_sleep', 'stall_no_inst', 'stall_tex', 'stall_drain')
for ki, k in enumerate(kern):
    if only and ki not in only:
        continue
    h = {n: i for i, n in enumerate(k['hdr'])}
    tot = sum(int(r[h['# Samples']] or 0) for r in k['rows'])
    agg = {c: sum(int(r[h[c]] or 0) for r in k['rows']) for c in ST}
    print(f"[{ki}] {k['name'][:70]} instrs {len(k['rows'])} samples {tot}")
    print("   ", ", ".join(f"{c[6:]} {100 * v / tot:.1f}%" for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    order = sorted(range(len(k['rows'])), key=lambda i: -int(k['rows'][i][h['# Samples']] or 0))[:top_n]
    for i in order:
        r = k['rows'][i]
        n = int(r[h['# Samples']])
        main = sorted(((c[6:], int(r[h[c]] or 0)) for c in ST), key=lambda kv: -kv[1])[:2]
        print(f"{i:5d} {n:6d} {100 * n / tot:5.1f}%  {r[h['Source']][:64]:64s} {main}")
