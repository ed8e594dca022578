"""One launch of each streaming training kernel at its largest / typical layer shape (B=32), for `ncu --set full`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops
DEV = "cuda"; B = 32
def rnd(*s, dt=torch.bfloat16): return torch.randn(*s, device=DEV).to(dt)
names = []
def case(name, fn):
    for _ in range(2): fn()
    names.append(name)
# upsample+concat backward, up4 shape: dcat [B,128,256,16+32]
dcat = rnd(B, 128, 256, 48); accs = rnd(B, 128, 256, 16)
case("upcat_bwd up4", lambda: ops.upcat_bwd(dcat, 16, accs))
# depthwise weight gradient: f3 (144ch @ 64x128) and f8 (384 @ 16x32), staging given (kernel only)
for (nm, H, W, C) in (("f3", 64, 128, 144), ("f8", 16, 32, 384)):
    x = rnd(B, H, W, C); dz = rnd(B, H, W, C); acc = torch.zeros(ops.NSLOT, 9, C, device=DEV, dtype=torch.float64)
    case(f"dw_wgrad {nm}", lambda x=x, dz=dz, acc=acc: ops.dw_wgrad(x, dz, 1, acc=acc))
    wb = rnd(9, C); bias = torch.zeros(C, device=DEV)
    case(f"dwconv_bf16w {nm}", lambda x=x, wb=wb, bias=bias: ops.dwconv3x3_bf16w(x, wb, bias, 1, 0))
# BatchNorm passes on the f2-expand shape (96ch @ 128x256) and f3-expand (144ch @ 64x128)
for (nm, H, W, C) in (("f2e", 128, 256, 96), ("f3e", 64, 128, 144)):
    z = rnd(B, H, W, C); da = rnd(B, H, W, C)
    g, b_ = torch.ones(C, device=DEV), torch.zeros(C, device=DEV); rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    case(f"bn_fwd {nm} (stats, finalize, apply)", lambda z=z: ops.bn_train_forward(z, g, b_, rm, rv, 1e-5, 0.1, 2))
    a, sv = ops.bn_train_forward(z, g, b_, rm, rv, 1e-5, 0.1, 2)
    case(f"bn_bwd {nm} (reduce, f64->f32, apply)", lambda da=da, z=z, sv=sv: ops.bn_train_backward(da, z, sv, 2))
# stem weight gradient
x0 = torch.randn(B, 3, 256, 512, device=DEV); dz0 = rnd(B, 128, 256, 32); dw0 = torch.zeros(3, 3, 3, 32, device=DEV)
case("smallcin_wgrad stem", lambda: ops.smallcin_wgrad(x0, dz0, 2, dw=dw0))
torch.cuda.synchronize(); print("CASES " + "|".join(names))
