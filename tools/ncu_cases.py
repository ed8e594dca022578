"""A handful of single launches at real layer shapes (B=16 to keep ncu replay short) for `ncu --set full`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops

DEV = "cuda"
B = int(os.environ.get("KB_BATCH", "16"))
which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["dw", "upcat", "final", "stem", "tc"]


def rnd(*shape, dt=torch.bfloat16):
    return torch.randn(*shape, device=DEV, dtype=torch.float32).to(dt)


def twice(fn):
    fn(); fn()
    torch.cuda.synchronize()


if "dw" in which:
    x = rnd(B, 64, 128, 144); w = rnd(9, 144, dt=torch.float32); b = rnd(144, dt=torch.float32)
    twice(lambda: ops.dwconv3x3(x, w, b, 1, 2))
    x = rnd(B, 128, 256, 96)
    w = rnd(9, 96, dt=torch.float32); b = rnd(96, dt=torch.float32)
    twice(lambda: ops.dwconv3x3(x, w, b, 2, 2))
if "upcat" in which:
    x = rnd(B, 64, 128, 64); sk = rnd(B, 128, 256, 16)
    twice(lambda: ops.upsample2x_concat(sk, x))
if "final" in which:
    lg = rnd(B, 128, 256, 16)
    twice(lambda: ops.upsample2x_ac_nchw(lg, 10, torch.bfloat16))
    twice(lambda: ops.upsample2x_ac_argmax(lg, 10))
if "stem" in which:
    x = torch.randn(B, 3, 256, 512, device=DEV).bfloat16()
    w = rnd(3, 3, 3, 32, dt=torch.float32); b = rnd(32, dt=torch.float32)
    twice(lambda: ops.conv3x3_smallcin(x, w, b, 2, 2, torch.bfloat16))
if "tc" in which:
    fl = int(os.environ.get("TC_FLAGS", "0"))
    for (H, W, Cin, Cout, taps) in ((128, 256, 32, 16, 1), (128, 256, 16, 96, 1), (128, 256, 32, 32, 9), (64, 128, 152, 64, 9), (16, 32, 1344, 256, 9)):
        x = rnd(B, H, W, Cin); w = rnd(Cout, taps * Cin) * 0.05; b = rnd(Cout, dt=torch.float32)
        twice(lambda: ops.conv_tc(x, w, b, taps, 1, flags=fl))
print("ncu_cases done")
