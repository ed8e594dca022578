"""Fused inverted-residual kernel at the f2 / f3 / f8 shapes (B=64) for an ncu capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops
B = int(os.environ.get("KB_BATCH", "64"))
def rnd(*s, dt=torch.bfloat16, k=1.0): return (torch.randn(*s, device="cuda") * k).to(dt)
for (H, W, Cin, Ce, Cout, st) in [(128, 256, 16, 96, 24, 2), (64, 128, 24, 144, 24, 1), (16, 32, 64, 384, 64, 1)]:
    x = rnd(B, H, W, Cin)
    we, be = rnd(Ce, Cin, k=0.2), ops.pad_channels(rnd(Ce, dt=torch.float32, k=0.1), 64)
    wd, bd = ops.pad_channels(rnd(9, Ce, k=0.3), 64), ops.pad_channels(rnd(Ce, dt=torch.float32, k=0.1), 64)
    wp, bp = rnd(Cout, Ce, k=0.1), ops.pad_channels(rnd(Cout, dt=torch.float32, k=0.1), 16)
    for _ in range(2):
        ops.mbconv(x, we, be, wd, bd, wp, bp, st, st == 1 and Cin == Cout)
torch.cuda.synchronize(); print("done")
