"""Micro-benchmarks of the training-only kernels at real layer shapes (B=32, 256x512 input)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops
DEV = "cuda"; B = int(os.environ.get("KB_BATCH", "32"))
flt = sys.argv[1] if len(sys.argv) > 1 else ""
def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
def rnd(*s, dt=torch.bfloat16): return torch.randn(*s, device=DEV).to(dt)
def rep(name, us, nbytes): print(f"{name:56s} {us:9.1f} us {nbytes/us/1e3:8.0f} GB/s", flush=True)
for (nm, H, W, C) in (("f1.dw/32ch", 128, 256, 32), ("f2.e/96ch", 128, 256, 96), ("f3.e/144ch", 64, 128, 144), ("up4/32ch", 128, 256, 32),
                      ("f8.e/384ch", 16, 32, 384), ("f18/1280ch", 8, 16, 1280), ("outc/16ch", 128, 256, 16)):
    z = rnd(B, H, W, C); da = rnd(B, H, W, C)
    g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV); rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    nb = z.numel() * 2
    if not flt or "bnf" in flt:
        us = timeit(lambda: ops.bn_train_forward(z, g, b, rm, rv, 1e-5, 0.1, 2)); rep(f"bn_train_forward (stats+finalize+apply) {nm}", us, 3 * nb)
    a, sv = ops.bn_train_forward(z, g, b, rm, rv, 1e-5, 0.1, 2)
    if not flt or "bnb" in flt:
        us = timeit(lambda: ops.bn_train_backward(da, z, sv, 2)); rep(f"bn_train_backward (reduce+apply) {nm}", us, 5 * nb)
for (nm, H, W, C, s_) in (("f1.dw", 128, 256, 32, 1), ("f2.dw", 128, 256, 96, 2), ("f3.dw", 64, 128, 144, 1), ("f8.dw", 16, 32, 384, 1)):
    if flt and "dw" not in flt: continue
    x = rnd(B, H, W, C); dz = rnd(B, (H - 1) // s_ + 1, (W - 1) // s_ + 1, C); w9 = rnd(9, C, dt=torch.float32)
    us = timeit(lambda: ops.dw_wgrad(x, dz, s_)); rep(f"dw_wgrad {nm} C={C} s{s_}", us, (x.numel() + dz.numel()) * 2)
    us = timeit(lambda: ops.dw_dgrad(dz, w9, tuple(x.shape), s_)); rep(f"dw_dgrad {nm} C={C} s{s_}", us, (x.numel() + dz.numel()) * 2)
