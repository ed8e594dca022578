"""Data gradient of up4.conv.0 (32 -> 80 channels, 3x3, 128x256, B=32): row-stacked kernel vs the tap-by-tap conv_tc."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops
B, H, W, Cin, Cout = 32, 128, 256, 32, 80
x = torch.randn(B, H, W, Cin, device="cuda").bfloat16(); w = (torch.randn(Cout, 9 * Cin, device="cuda") * 0.05).bfloat16()
res = torch.randn(B, H, W, Cout, device="cuda").bfloat16()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
out = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
for nm, r in (("no res", None), ("res", res)):
    t_rs = timeit(lambda: ops.conv_rs(x, w, None, 0, r, out))
    t_tc = timeit(lambda: ops.conv_tc(x, w, None, 9, 0, r, flags=2))
    a = ops.conv_rs(x, w, None, 0, r); b = ops.conv_tc(x, w, None, 9, 0, r, flags=2)
    print(f"{nm}: conv_rs<80> {t_rs:.1f} us   conv_tc {t_tc:.1f} us   max diff {float((a.float() - b.float()).abs().max()):.3e}")
