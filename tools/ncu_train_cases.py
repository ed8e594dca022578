import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops
DEV = "cuda"; B = 32
def rnd(*s, dt=torch.bfloat16): return torch.randn(*s, device=DEV).to(dt)
x = rnd(B, 64, 128, 144); dz = rnd(B, 64, 128, 144); w9 = rnd(9, 144, dt=torch.float32)
for _ in range(2):
    ops.dw_wgrad(x, dz, 1); ops.dw_dgrad(dz, w9, tuple(x.shape), 1)
z = rnd(B, 128, 256, 96); da = rnd(B, 128, 256, 96)
g, b = torch.ones(96, device=DEV), torch.zeros(96, device=DEV); rm, rv = torch.zeros(96, device=DEV), torch.ones(96, device=DEV)
for _ in range(2):
    a, sv = ops.bn_train_forward(z, g, b, rm, rv, 1e-5, 0.1, 2); ops.bn_train_backward(da, z, sv, 2)
torch.cuda.synchronize(); print("done")
