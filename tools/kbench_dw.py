"""Depthwise 3x3 kernels at the MobileNetV2 layer shapes: f32-tap SIMT kernel vs the bf16-tap mixed-precision-FMA kernel
(variants = register blockings).  KB_BATCH=64 (eval) / 32 (training).  Checks the two agree bit for bit on bf16 taps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops
from kbench import timeit

B = int(os.environ.get("KB_BATCH", "64"))
LAYERS = [("f1", 128, 256, 32, 1), ("f2", 128, 256, 96, 2), ("f3", 64, 128, 144, 1), ("f4", 64, 128, 144, 2), ("f5", 32, 64, 192, 1),
          ("f7", 32, 64, 192, 2), ("f8", 16, 32, 384, 1), ("f11", 16, 32, 384, 1), ("f12", 16, 32, 576, 1), ("f14", 16, 32, 576, 2),
          ("f15", 8, 16, 960, 1)]
for name, H, W, C, s in LAYERS:
    x = torch.randn(B, H, W, C, device="cuda").bfloat16()
    w = (torch.randn(9, C, device="cuda") * 0.3).bfloat16()
    wf = w.float().contiguous()
    b = torch.randn(C, device="cuda")
    Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
    nbytes = (x.numel() + B * Ho * Wo * C) * 2
    y0 = ops.dwconv3x3(x, wf, b, s, 2)
    t0 = timeit(lambda: ops.dwconv3x3(x, wf, b, s, 2, out=y0))
    line = f"{name:4s} C={C:4d} s{s} @{H}x{W}  f32-tap {t0:7.1f} us ({nbytes / t0 / 1e3:5.0f} GB/s) |"
    for v in (0, 1, 2, 3):
        y = ops.dwconv3x3_bf16w(x, w, b, s, 2, variant=v)
        same = bool(torch.equal(y, y0))
        t = timeit(lambda: ops.dwconv3x3_bf16w(x, w, b, s, 2, out=y, variant=v))
        line += f" v{v} {t:7.1f} us ({nbytes / t / 1e3:5.0f} GB/s){'' if same else ' MISMATCH'}"
    print(line, flush=True)
