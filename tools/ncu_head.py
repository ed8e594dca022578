"""Fused head kernel (stem_mb1) at the bench shape (B=64, 3x256x512 bf16 frames) for an ncu capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops
B = int(os.environ.get("KB_BATCH", "64"))
g = torch.Generator(device="cpu").manual_seed(0)
x = torch.randn(B, 3, 256, 512, generator=g).bfloat16().cuda()
w0 = (torch.randn(3, 3, 3, 32, generator=g) * 0.27).cuda(); b0 = (torch.randn(32, generator=g) * 0.3).cuda()
wd = (torch.randn(9, 32, generator=g) * 0.4).bfloat16().cuda(); bd = (torch.randn(32, generator=g) * 0.3).cuda()
wp = (torch.randn(16, 32, generator=g) * 0.18).bfloat16().cuda(); bp = (torch.randn(16, generator=g) * 0.2).cuda()
y = torch.empty(B, 128, 256, 16, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.stem_mb1(x, w0, b0, wd, bd, wp, bp, out=y)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.stem_mb1(x, w0, b0, wd, bd, wp, bp, out=y)
e1.record(); torch.cuda.synchronize()
print(f"stem_mb1 B={B}: {e0.elapsed_time(e1) * 100:.1f} us")
