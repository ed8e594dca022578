"""How many global loads each kernel keeps in flight: per kernel of an object file / library, the sizes of the LDG
clusters (loads separated by <= GAP other instructions) in its SASS.  A streaming kernel whose largest cluster is 1-2
is latency-bound whatever its unroll pragma says (ptxas sinks loads into the arithmetic)."""
import re, subprocess, sys
GAP = 6
obj = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
name = None; instrs = []
def flush():
    if name is None or (flt and not re.search(flt, name)): return
    idx = [i for i, s in enumerate(instrs) if s.startswith(("LDG", "LD.", "LDGSTS", "@")) and "LDG" in s]
    if not idx: return
    clusters = []; cur = 1
    for a, b in zip(idx, idx[1:]):
        if b - a <= GAP: cur += 1
        else: clusters.append(cur); cur = 1
    clusters.append(cur)
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:110]
    print(f"{len(instrs):6d} instr  {len(idx):4d} LDG  max cluster {max(clusters):3d}  clusters {sorted(clusters, reverse=True)[:8]}  {short}")
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: flush(); name = m.group(1); instrs = []; continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m: instrs.append(m.group(1).strip())
flush()
