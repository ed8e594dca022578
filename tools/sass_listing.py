"""Trimmed SASS evidence of libb200seg.so: per kernel, the count of the Blackwell mnemonics that prove the tcgen05 / TMEM / TMA
path (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit), of the
mixed-precision FMA (FHFMA.BF16), of legacy tensor-core instructions (HMMA = mma.sync), cluster instructions (UCGABAR /
barrier.cluster, remote shared-memory loads) and async copies (LDGSTS = cp.async); plus registers per thread.
    python tools/sass_listing.py > profiles/r02_sass_listing.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "team02-objectdetection_b200", "b200seg", "libb200seg.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
regs = {}
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m: cur = m.group(1); continue
    m = re.search(r"REG:(\d+)", line)
    if m and cur: regs[cur] = int(m.group(1)); cur = None
MN = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "FHFMA", "HMMA", "UCGABAR", "LDGSTS", "REDG", "RED.", "ATOMG"]
rows = []
name, cnt, n = None, None, 0
def flush():
    if name is not None: rows.append((name, n, dict(cnt)))
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush(); name, cnt, n = m.group(1), collections.Counter(), 0; continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?(\S+)", line)
    if m and name is not None:
        n += 1
        op = m.group(1)
        for k in MN:
            if op.startswith(k): cnt[k] += 1
flush()
dem = subprocess.run(["c++filt"], input="\n".join(r[0] for r in rows), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.relpath(lib, ROOT)}: {len(rows)} sm_100a kernels; totals: " +
      ", ".join(f"{k} {sum(r[2].get(k, 0) for r in rows)}" for k in MN if sum(r[2].get(k, 0) for r in rows)))
print(f"{'kernel':86s} {'instr':>6s} {'regs':>4s}  mnemonic counts")
for (nm, n, c), d in sorted(zip(rows, dem), key=lambda t: t[1]):
    short = re.sub(r"\(.*", "", d.replace("void ", "").replace("b200::", "").replace("(anonymous namespace)::", "").replace("__nv_bfloat16", "bf16"))[:86]
    tags = " ".join(f"{k}={v}" for k, v in c.items() if v)
    print(f"{short:86s} {n:6d} {regs.get(nm, 0):4d}  {tags}")
