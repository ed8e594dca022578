"""Row-stacked 3x3 conv (conv_rs) against conv_tc at the up4 shapes (forward B=64, and the training shapes B=32)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops
from kbench import timeit
for B in (64, 32):
    for name, Cin, Cout in (("up4.conv.0", 80, 32), ("up4.conv.3", 32, 32), ("up4.conv.0 dgrad-like", 32, 32)):
        H, W = 128, 256
        x = torch.randn(B, H, W, Cin, device="cuda").bfloat16()
        w = (torch.randn(Cout, 9 * Cin, device="cuda") * 0.05).bfloat16()
        b = torch.randn(Cout, device="cuda")
        y = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
        nbytes = (x.numel() + y.numel() + w.numel()) * 2
        t0 = timeit(lambda: ops.conv_tc(x, w, b, 9, 1, None, out=y, flags=1 << 30))
        line = f"B={B} {name:22s} {Cin}->{Cout}  conv_tc {t0:7.1f} us ({nbytes / t0 / 1e3:5.0f} GB/s) | conv_rs"
        for fl, tag in ((0, "auto"), (1 << 20, "1cta"), (2 << 20, "2cta"), ((4 << 16) | (1 << 20), "1cta/4st"), ((6 << 16) | (1 << 20), "1cta/6st"), ((3 << 16) | (2 << 20), "2cta/3st")):
            try:
                t = timeit(lambda: ops.conv_rs(x, w, b, 1, None, out=y, flags=fl))
                line += f"  {tag} {t:6.1f} ({nbytes / t / 1e3:4.0f})"
            except RuntimeError as e:
                line += f"  {tag} n/a"
        print(line, flush=True)
