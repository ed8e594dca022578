"""CTAs-per-SM / ring-depth sweep for the small-N decoder convs (DESIGN.md finding 8: they sit at ~50 % of the MMA issue floor)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops
B = 64
def rnd(*s, dt=torch.bfloat16): return torch.randn(*s, device="cuda").to(dt)
def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for name, H, W, Cin, Cout, taps in [("up4.conv.3", 128, 256, 32, 32, 9), ("up4.conv.0", 128, 256, 80, 32, 9), ("outc.conv.0", 128, 256, 32, 32, 1),
                                    ("up3.conv.3", 64, 128, 64, 64, 9), ("up3.conv.0", 64, 128, 152, 64, 9), ("f1.conv.1", 128, 256, 32, 16, 1)]:
    x = rnd(B, H, W, Cin); w = rnd(Cout, taps * Cin) * 0.05; b = rnd(Cout, dt=torch.float32)
    out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
    line = f"{name:12s} {Cin:3d}->{Cout:3d} k{taps} @{H}x{W}:"
    for ps in (0, 1, 2, 3):
        for st in (0, 2, 3, 4):
            fl = (st << 16) | (ps << 20)
            try:
                us = timeit(lambda: ops.conv_tc(x, w, b, taps, 1, None, out=out, flags=fl))
                line += f"  ps{ps}/st{st} {us:6.1f}"
            except RuntimeError:
                line += f"  ps{ps}/st{st}    n/a"
    print(line, flush=True)
