"""One launch of each training kernel rewritten at the end of round 2, at its layer shape of the B=32 step, for `ncu --set full`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops
DEV = "cuda"; B = 32
def rnd(*s, dt=torch.bfloat16): return torch.randn(*s, device=DEV).to(dt)
# cross-entropy fwd+grad on the step's logits
lg = torch.randn(B, 10, 256, 512, device=DEV); tg = torch.randint(0, 10, (B, 256, 512), device=DEV)
for _ in range(2): ops.softmax_ce(lg, tg)
# stem weight gradient
x0 = torch.randn(B, 3, 256, 512, device=DEV); dz0 = rnd(B, 128, 256, 32); dw0 = torch.zeros(3, 3, 3, 32, device=DEV)
for _ in range(2): ops.smallcin_wgrad(x0, dz0, 2, dw=dw0)
# large-layer BatchNorm with the constants derived in the apply kernels (f2 expand: 96 ch @ 128x256)
C = 96; z = rnd(B, 128, 256, C); da = rnd(B, 128, 256, C)
g, b_ = torch.ones(C, device=DEV), torch.zeros(C, device=DEV); rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
red = torch.zeros(ops.NSLOT, 2, C, device=DEV, dtype=torch.float64)
for _ in range(2):
    a, sv = ops.bn_train_forward(z, g, b_, rm, rv, 1e-5, 0.1, 2)
    red.zero_(); ops.bn_train_backward(da, z, sv, 2, red=red)
# dense weight gradient of up4.conv.3 (M = 64) and the 80-column row-stacked data gradient of up4.conv.0
x = rnd(B, 128, 256, 32); dz = rnd(B, 128, 256, 32); dw = torch.zeros(32, 9 * 32, device=DEV)
w80 = (torch.randn(80, 9 * 32, device=DEV) * 0.05).bfloat16()
for _ in range(2):
    ops.conv_wgrad_tc(x, dz, 9, dw=dw); ops.conv_rs(dz, w80, None, 0)
torch.cuda.synchronize(); print("done")
