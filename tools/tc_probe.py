"""GPU probe for the tcgen05 conv kernel: runs cases from simple to complex, prints error structure.
Usage (on a GPU box): python tools/tc_probe.py [direct]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
import torch.nn.functional as F
from b200seg import ops

DEV = "cuda"
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 0
CASES = [  # B,H,W,Cin,Cout,taps,act,res
    (1, 8, 16, 64, 64, 1, 0, 0), (1, 8, 16, 64, 16, 1, 0, 0), (1, 8, 16, 16, 64, 1, 0, 0), (1, 8, 16, 128, 64, 1, 0, 0),
    (1, 8, 16, 512, 64, 1, 0, 0), (1, 8, 16, 64, 256, 1, 0, 0), (2, 16, 32, 64, 64, 1, 2, 1), (1, 8, 16, 64, 384, 1, 0, 0),
    (1, 7, 9, 40, 72, 1, 2, 1), (1, 8, 16, 64, 64, 9, 0, 0), (2, 16, 32, 64, 64, 9, 1, 0), (1, 23, 40, 64, 64, 9, 1, 0),
    (2, 16, 32, 1344, 256, 9, 1, 0), (1, 128, 256, 80, 32, 9, 1, 0), (2, 64, 128, 152, 64, 9, 1, 0), (1, 33, 100, 32, 32, 9, 1, 0),
    (3, 64, 128, 64, 64, 9, 1, 0), (2, 128, 256, 32, 32, 9, 1, 0), (4, 64, 128, 16, 96, 1, 2, 0), (4, 64, 128, 144, 24, 1, 0, 1),
]
for (B, H, W, Cin, Cout, taps, act, res) in CASES:
    k = 3 if taps == 9 else 1
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, H, W, generator=g).to(DEV).bfloat16()
    w = (torch.randn(Cout, Cin, k, k, generator=g) * (2.0 / (Cin * taps)) ** 0.5).to(DEV).bfloat16()
    b = (torch.randn(Cout, generator=g) * 0.1).to(DEV)
    r = torch.randn(B, Cout, H, W, generator=g).to(DEV).bfloat16() if res else None
    ref = F.conv2d(x.double(), w.double(), b.double(), 1, k // 2)
    ref = {0: ref, 1: F.relu(ref), 2: ref.clamp(0, 6)}[act]
    if res:
        ref = ref + r.double()
    try:
        got = ops.conv_tc(x.permute(0, 2, 3, 1).contiguous(), w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous(), b, taps, act,
                          r.permute(0, 2, 3, 1).contiguous() if res else None, flags=flags)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print("CASE", (B, H, W, Cin, Cout, taps, act, res), "EXCEPTION", repr(e)[:300]); break
    got = got.permute(0, 3, 1, 2).double()
    d = (got - ref).abs()
    e = float(d.max() / ref.abs().max())
    print("CASE", (B, H, W, Cin, Cout, taps, act, res), f"flags={flags} err={e:.3e}", "OK" if e < 6e-3 else "BAD")
    if e >= 6e-3:
        bad = d > 6e-3 * ref.abs().max()
        print("   bad frac", float(bad.float().mean()), "nan", int(torch.isnan(got).sum()), "zeros", float((got == 0).float().mean()))
        print("   bad by channel (first 32):", bad.float().mean(dim=(0, 2, 3))[:32].tolist())
        pix = bad.float().mean(dim=1).flatten()
        print("   bad by pixel (first 48):", [round(v, 2) for v in pix[:48].tolist()])
        print("   got[0,:4,0,:4]", got[0, :4, 0, :4].tolist()); print("   ref[0,:4,0,:4]", ref[0, :4, 0, :4].tolist())
        # is it a K-subset? compare with the conv using only the first 16/32/48 input channels
        if taps == 1:
            for kk in (16, 32, 48, 64):
                if kk <= Cin:
                    part = F.conv2d(x[:, :kk].double(), w[:, :kk].double(), b.double())
                    print(f"   err vs K[:{kk}] partial:", float((got - part).abs().max() / ref.abs().max()))
