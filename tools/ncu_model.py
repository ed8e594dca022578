"""One eager (graph-free) bf16 eval forward of MobileNetV2UNet at the bench shape (B=64, 3x256x512) so that `ncu --set full`
sees every kernel of the schedule once with its real operands.  Prints the schedule (launch order) for the summary."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
import b200seg
B = int(os.environ.get("KB_BATCH", "64"))
torch.manual_seed(0)
m = b200seg.MobileNetV2UNet(output_channels=10).cuda().eval()
eng = m._get_engine(); eng.precision = "bf16"; eng.use_graphs = False
x = torch.randn(B, 3, 256, 512, device="cuda").bfloat16()
with torch.no_grad():
    for it in range(int(os.environ.get("NCU_FWD", "2"))):
        y = m(x)
torch.cuda.synchronize()
print("SCHEDULE " + "|".join(f"{s.op}:{s.name}" for s in eng._schedule("bf16", "tc", 256, 512, torch.bfloat16)))
