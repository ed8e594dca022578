"""conv_wgrad_tc at the decoder / encoder layer shapes of the training step (B=32): kernel-only time (staging given),
CUDA events over 20 back-to-back launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops
B = 32
SHAPES = [("outc.conv.0", 32, 16, 1, 128, 256), ("up4.conv.3", 32, 32, 9, 128, 256), ("up4.conv.0", 80, 32, 9, 128, 256),
          ("up3.conv.3", 64, 64, 9, 64, 128), ("up3.conv.0", 152, 64, 9, 64, 128), ("up2.conv.0", 288, 128, 9, 32, 64),
          ("up1.conv.0", 1344, 256, 9, 16, 32), ("f17.conv.2", 960, 320, 1, 8, 16), ("f8.conv.0", 64, 384, 1, 16, 32),
          ("f3.conv.0", 24, 144, 1, 64, 128), ("f2.conv.0", 16, 96, 1, 128, 256)]
for (nm, cin, cout, taps, H, W) in SHAPES:
    x = torch.randn(B, H, W, cin, device="cuda").bfloat16(); dz = torch.randn(B, H, W, cout, device="cuda").bfloat16()
    dw = torch.zeros(cout, taps * cin, device="cuda")
    for _ in range(3): ops.conv_wgrad_tc(x, dz, taps, dw=dw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.conv_wgrad_tc(x, dz, taps, dw=dw)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    # correctness spot check against an fp32 matmul on one tap (centre)
    dw.zero_(); ops.conv_wgrad_tc(x, dz, taps, dw=dw)
    ref = dz.float().reshape(-1, cout).t() @ x.float().reshape(-1, cin)
    got = dw.view(cout, taps, cin)[:, taps // 2, :]
    err = float((got - ref).abs().max() / ref.abs().max())
    print(f"{nm:12s} cin {cin:4d} cout {cout:4d} taps {taps} {H}x{W}  {us:8.1f} us   centre-tap err {err:.2e}", flush=True)
