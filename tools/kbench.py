"""Micro-benchmarks of single kernels at the real MobileNetV2UNet layer shapes (B=64, 256x512 input).
Usage: python tools/kbench.py [filter]   -- prints us, GB/s (algorithmic), TF/s per case."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops

DEV = "cuda"
B = int(os.environ.get("KB_BATCH", "64"))
flt = sys.argv[1] if len(sys.argv) > 1 else ""


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
    return e0.elapsed_time(e1) / iters * 1e3  # us


def report(name, us, nbytes, flops=0):
    print(f"{name:58s} {us:9.1f} us  {nbytes / us / 1e3:8.0f} GB/s  {flops / us / 1e6:8.1f} TF/s", flush=True)


def rnd(*shape, dt=torch.bfloat16):
    return torch.randn(*shape, device=DEV, dtype=torch.float32).to(dt)


def conv_case(name, H, W, Cin, Cout, taps, res=False, flag_list=(0,)):
    x = rnd(B, H, W, Cin); w = rnd(Cout, taps * Cin) * 0.05; b = rnd(Cout, dt=torch.float32)
    r = rnd(B, H, W, Cout) if res else None
    out = torch.empty(B, H, W, Cout, device=DEV, dtype=torch.bfloat16)
    nbytes = (x.numel() + out.numel() + w.numel() + (r.numel() if res else 0)) * 2
    flops = 2 * B * H * W * Cout * taps * Cin
    for fl in flag_list:
        nm = f"conv_tc {name} {Cin}->{Cout} k{taps} @{H}x{W} flags={fl} (res={1 - ((fl >> 2) & 1)} st={(fl >> 16) & 15} ps={(fl >> 20) & 3})"
        if flt and flt not in nm:
            continue
        us = timeit(lambda: ops.conv_tc(x, w, b, taps, 1, r, out=out, flags=fl))
        report(nm, us, nbytes, flops)


def main():
    def F(res=1, stages=0, per_sm=0, halo=1):
        return (0 if res else 4) | (0 if halo else 2) | (stages << 16) | (per_sm << 20)
    if os.environ.get("KB_SWEEP"):
        sweep = [F(r, st, ps) for r in (1, 0) for st in (2, 3, 4, 6) for ps in (0,)]
        conv_case("up4.0", 128, 256, 80, 32, 9, flag_list=sweep)
        conv_case("up4.3", 128, 256, 32, 32, 9, flag_list=sweep)
        conv_case("up3.0", 64, 128, 152, 64, 9, flag_list=sweep)
        conv_case("up3.3", 64, 128, 64, 64, 9, flag_list=sweep)
        conv_case("up2.0", 32, 64, 288, 128, 9, flag_list=[F(1, st) for st in (3, 4, 6, 8)])
        conv_case("up1.0", 16, 32, 1344, 256, 9, flag_list=[F(1, st) for st in (2, 3, 4)])
        sweep1 = [F(r, st, ps) for r in (1, 0) for st in (2, 4, 6) for ps in (0, 1)]
        conv_case("f2.expand", 128, 256, 16, 96, 1, flag_list=sweep1)
        conv_case("f1.project", 128, 256, 32, 16, 1, flag_list=sweep1)
        conv_case("f3.expand", 64, 128, 24, 144, 1, flag_list=sweep1)
        conv_case("f5.expand", 32, 64, 32, 192, 1, flag_list=sweep1)
        conv_case("f8.expand", 16, 32, 64, 384, 1, flag_list=[F(0, st) for st in (2, 4, 6)])
        conv_case("f18", 8, 16, 320, 1280, 1, flag_list=[F(0, st) for st in (2, 3, 4)])
        for (nm, H, W, C, s_) in (("f1.dw", 128, 256, 32, 1), ("f2.dw", 128, 256, 96, 2), ("f3.dw", 64, 128, 144, 1), ("f5.dw", 32, 64, 192, 1), ("f8.dw", 16, 32, 384, 1)):
            x = rnd(B, H, W, C); w = rnd(9, C, dt=torch.float32); b = rnd(C, dt=torch.float32)
            out = torch.empty(B, (H - 1) // s_ + 1, (W - 1) // s_ + 1, C, device=DEV, dtype=torch.bfloat16)
            wd = ops.pack_dw_diag(w)
            for fl in [F(r, st, ps) for r in (1, 0) for st in (2, 4, 6) for ps in (0, 1)]:
                us = timeit(lambda: ops.dwconv3x3_tc(x, wd, b, s_, 2, out=out, flags=fl))
                report(f"dw_tc {nm} C={C} s{s_} flags={fl} (res={1 - ((fl >> 2) & 1)} st={(fl >> 16) & 15} ps={(fl >> 20) & 3})", us, (x.numel() + out.numel()) * 2)
        return
    FL3 = (0, 32, 2)                 # default | no M-tile pairing | no halo
    conv_case("up4.0", 128, 256, 80, 32, 9, flag_list=FL3)
    conv_case("up4.3", 128, 256, 32, 32, 9, flag_list=FL3)
    conv_case("up3.0", 64, 128, 152, 64, 9, flag_list=FL3)
    conv_case("up3.3", 64, 128, 64, 64, 9, flag_list=FL3)
    conv_case("up2.0", 32, 64, 288, 128, 9, flag_list=(0, 32))
    conv_case("up2.3", 32, 64, 128, 128, 9, flag_list=(0, 32))
    conv_case("up1.0", 16, 32, 1344, 256, 9, flag_list=(0, 32))
    conv_case("up1.3", 16, 32, 256, 256, 9, flag_list=(0, 32))
    conv_case("f2.expand", 128, 256, 16, 96, 1)
    conv_case("f1.project", 128, 256, 32, 16, 1)
    conv_case("f2.project", 64, 128, 96, 24, 1)
    conv_case("f3.expand", 64, 128, 24, 144, 1)
    conv_case("f3.project", 64, 128, 144, 24, 1, res=True)
    conv_case("f5.expand", 32, 64, 32, 192, 1)
    conv_case("f8.expand", 16, 32, 64, 384, 1)
    conv_case("f8.project", 16, 32, 384, 64, 1, res=True)
    conv_case("f15.expand", 8, 16, 160, 960, 1)
    conv_case("f17.project", 8, 16, 960, 320, 1)
    conv_case("f18", 8, 16, 320, 1280, 1)
    conv_case("outc.0", 128, 256, 32, 16, 1)
    conv_case("outc.3", 128, 256, 16, 16, 1)
    # HBM ops
    for (nm, H, W, C, s) in (("f1.dw", 128, 256, 32, 1), ("f2.dw", 128, 256, 96, 2), ("f3.dw", 64, 128, 144, 1),
                             ("f4.dw", 64, 128, 144, 2), ("f5.dw", 32, 64, 192, 1), ("f8.dw", 16, 32, 384, 1), ("f15.dw", 8, 16, 960, 1)):
        name = f"dwconv {nm} C={C} s{s} @{H}x{W}"
        if flt and flt not in name:
            continue
        x = rnd(B, H, W, C); w = rnd(9, C, dt=torch.float32); b = rnd(C, dt=torch.float32)
        out = torch.empty(B, (H - 1) // s + 1, (W - 1) // s + 1, C, device=DEV, dtype=torch.bfloat16)
        us = timeit(lambda: ops.dwconv3x3(x, w, b, s, 2, out=out))
        report(name, us, (x.numel() + out.numel()) * 2, 18 * out.numel())
        wd = ops.pack_dw_diag(w)
        for fl in (0, 2):
            us = timeit(lambda: ops.dwconv3x3_tc(x, wd, b, s, 2, out=out, flags=fl))
            report(name + f" TENSOR-CORE flags={fl}", us, (x.numel() + out.numel()) * 2, 18 * out.numel())
    for (nm, h, w_, Cs, Cu) in (("up1", 8, 16, 64, 1280), ("up2", 16, 32, 32, 256), ("up3", 32, 64, 24, 128), ("up4", 64, 128, 16, 64)):
        name = f"upcat {nm} {Cs}+{Cu} @{h}x{w_}"
        if flt and flt not in name:
            continue
        x = rnd(B, h, w_, Cu); sk = rnd(B, 2 * h, 2 * w_, Cs)
        out = torch.empty(B, 2 * h, 2 * w_, Cs + Cu, device=DEV, dtype=torch.bfloat16)
        us = timeit(lambda: ops.upsample2x_concat(sk, x, out=out))
        report(name, us, (x.numel() + sk.numel() + out.numel()) * 2)
    if not flt or "final" in flt:
        lg = rnd(B, 128, 256, 16)
        out = torch.empty(B, 10, 256, 512, device=DEV, dtype=torch.bfloat16)
        us = timeit(lambda: ops.upsample2x_ac_nchw(lg, 10, torch.bfloat16, out=out))
        report("final upsample -> NCHW bf16", us, (lg.numel() + out.numel()) * 2)
        mk = torch.empty(B, 256, 512, device=DEV, dtype=torch.uint8)
        us = timeit(lambda: ops.upsample2x_ac_argmax(lg, 10, out=mk))
        report("final upsample -> argmax mask", us, lg.numel() * 2 + mk.numel())
    if not flt or "stem" in flt:
        for dt in (torch.bfloat16, torch.float32):
            x = torch.randn(B, 3, 256, 512, device=DEV).to(dt)
            w = rnd(3, 3, 3, 32, dt=torch.float32); b = rnd(32, dt=torch.float32)
            out = torch.empty(B, 128, 256, 32, device=DEV, dtype=torch.bfloat16)
            us = timeit(lambda: ops.conv3x3_smallcin(x, w, b, 2, 2, torch.bfloat16, out=out))
            report(f"stem 3->32 s2 in={dt}", us, x.numel() * x.element_size() + out.numel() * 2, 2 * 27 * out.numel())


if __name__ == "__main__":
    main()
