"""How many global loads each kernel keeps in flight: per kernel of an object file / library, the sizes of the LDG
clusters (loads separated by <= GAP other instructions) in its SASS.  A streaming kernel whose largest cluster is 1-2
is latency-bound whatever its unroll pragma says (ptxas sinks loads into the arithmetic).

    python tools/sass_ldg_clusters.py OBJ_OR_SO [regex]          (also imported by tests/test_cabi.py)"""
import re, subprocess, sys
GAP = 6


def ldg_clusters(obj, flt="", wide_only=False):
    """{demangled kernel name: (instruction count, [LDG cluster sizes, largest first])}; wide_only: 16-byte loads only (the
    streamed activations -- per-channel constants are scalar loads)"""
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    kernels, name, instrs = [], None, []
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if name is not None: kernels.append((name, instrs))
            name, instrs = m.group(1), []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and name is not None: instrs.append(m.group(1).strip())
    if name is not None: kernels.append((name, instrs))
    dem = subprocess.run(["c++filt"], input="\n".join(k[0] for k in kernels), capture_output=True, text=True).stdout.splitlines()
    res = {}
    for (mangled, ins), d in zip(kernels, dem):
        if flt and not re.search(flt, mangled) and not re.search(flt, d): continue
        idx = [i for i, s in enumerate(ins) if ("LDG" in s.split(" ")[0] or (s.startswith("@") and "LDG" in s))
               and (not wide_only or ".128" in s)]
        if not idx: continue
        clusters, cur = [], 1
        for a, b in zip(idx, idx[1:]):
            if b - a <= GAP: cur += 1
            else: clusters.append(cur); cur = 1
        clusters.append(cur)
        res[d] = (len(ins), sorted(clusters, reverse=True))
    return res


if __name__ == "__main__":
    for d, (n, cl) in ldg_clusters(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "", wide_only="--wide" in sys.argv).items():
        print(f"{n:6d} instr  {sum(cl):4d} LDG  max cluster {cl[0]:3d}  clusters {cl[:8]}  {d[:110]}")
