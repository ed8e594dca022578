"""The bench's dominant kernel at the bench's size (B=64) for the roofline `traffic` field."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg import ops
B = 64
def rnd(*s, dt=torch.bfloat16): return torch.randn(*s, device="cuda").to(dt)
x = rnd(B, 128, 256, 80); w = rnd(32, 9 * 80) * 0.05; b = rnd(32, dt=torch.float32)
for _ in range(3): ops.conv_tc(x, w, b, 9, 1)
x2 = rnd(B, 128, 256, 96); w9 = rnd(9, 96, dt=torch.float32); b2 = rnd(96, dt=torch.float32)
for _ in range(3): ops.dwconv3x3(x2, w9, b2, 2, 2)
torch.cuda.synchronize(); print("done")
