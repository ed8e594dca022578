"""Train-mode BatchNorm passes on a mid-size layer (f8 expand: 384 ch @ 16x32, B=32 -> 12.6 MB per tensor) for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops
C, H, W, B = int(os.environ.get("BN_C", "384")), int(os.environ.get("BN_H", "16")), int(os.environ.get("BN_W", "32")), 32
z = torch.randn(B, H, W, C, device="cuda").bfloat16(); da = torch.randn(B, H, W, C, device="cuda").bfloat16()
g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda"); rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
for _ in range(3):
    a, sv = ops.bn_train_forward(z, g, b, rm, rv, 1e-5, 0.1, 2); ops.bn_train_backward(da, z, sv, 2)
torch.cuda.synchronize(); print("done")
