"""Fused inverted-residual kernel vs the three unfused kernels at the encoder's real shapes (B=64, 256x512 input).
Usage: python tools/kbench_mb.py [flags ...]  -- prints us per block for the unfused path and for each flags value."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops

DEV = "cuda"
B = int(os.environ.get("KB_BATCH", "64"))
FLAGS = [int(a, 0) for a in sys.argv[1:]] or [0]

BLOCKS = [  # name, H, W (input), Cin, Ce, Cout, stride
    ("f2", 128, 256, 16, 96, 24, 2), ("f3", 64, 128, 24, 144, 24, 1), ("f4", 64, 128, 24, 144, 32, 2),
    ("f5", 32, 64, 32, 192, 32, 1), ("f7", 32, 64, 32, 192, 64, 2), ("f8", 16, 32, 64, 384, 64, 1),
    ("f11", 16, 32, 64, 384, 96, 1), ("f12", 16, 32, 96, 576, 96, 1), ("f14", 16, 32, 96, 576, 160, 2),
    ("f15", 8, 16, 160, 960, 160, 1), ("f17", 8, 16, 160, 960, 320, 1),
]


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def rnd(*shape, dt=torch.bfloat16, s=1.0):
    return (torch.randn(*shape, device=DEV, dtype=torch.float32) * s).to(dt)


tot_u, tot_f = 0.0, {f: 0.0 for f in FLAGS}
for name, H, W, Cin, Ce, Cout, st in BLOCKS:
    res = st == 1 and Cin == Cout
    x = rnd(B, H, W, Cin)
    we, be = rnd(Ce, Cin, s=0.2), rnd(Ce, dt=torch.float32, s=0.1)
    wd, bd = rnd(9, Ce, dt=torch.float32, s=0.3), rnd(Ce, dt=torch.float32, s=0.1)
    wp, bp = rnd(Cout, Ce, s=0.1), rnd(Cout, dt=torch.float32, s=0.1)
    Ho, Wo = (H - 1) // st + 1, (W - 1) // st + 1
    e = torch.empty(B, H, W, Ce, device=DEV, dtype=torch.bfloat16)
    d = torch.empty(B, Ho, Wo, Ce, device=DEV, dtype=torch.bfloat16)
    y = torch.empty(B, Ho, Wo, Cout, device=DEV, dtype=torch.bfloat16)
    use_tc_dw = st == 1 and W >= 96
    wdiag = ops.pack_dw_diag(wd) if use_tc_dw else None

    def unfused():
        ops.conv_tc(x, we, be, 1, 2, None, out=e)
        if use_tc_dw:
            ops.dwconv3x3_tc(e, wdiag, bd, st, 2, out=d)
        else:
            ops.dwconv3x3(e, wd, bd, st, 2, out=d)
        ops.conv_tc(d, wp, bp, 1, 0, x if res else None, out=y)

    pe, pwd, pbd, pbp = ops.pad_channels(be, 64), ops.pad_channels(wd.bfloat16(), 64), ops.pad_channels(bd, 64), ops.pad_channels(bp, 16)
    y2 = torch.empty_like(y)
    us_u = timeit(unfused)
    tot_u += us_u
    io_bytes = (x.numel() + y.numel()) * 2
    line = f"{name:4s} {Cin:3d}->{Ce:3d}->{Cout:3d} s{st} @{H}x{W}  unfused {us_u:7.1f} us |"
    for f in FLAGS:
        try:
            us_f = timeit(lambda: ops.mbconv(x, we, pe, pwd, pbd, wp, pbp, st, res, out=y2, flags=f))
        except RuntimeError:
            line += f"  fused[{f:#x}]     n/a (does not fit)          "
            continue
        tot_f[f] += us_f
        line += f"  fused[{f:#x}] {us_f:7.1f} us ({io_bytes / us_f / 1e3:5.0f} GB/s io)"
    unfused()
    err = float((y2.float() - y.float()).abs().max() / y.float().abs().max())
    print(line + f"  maxdiff {err:.2e}", flush=True)
print(f"sum (one of each shape): unfused {tot_u:.0f} us; " + "; ".join(f"fused[{f:#x}] {v:.0f} us" for f, v in tot_f.items()))
