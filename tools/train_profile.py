"""Kernel-level time table of one eager training step (torch.profiler / CUPTI)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
import b200seg
from torch.profiler import profile, ProfilerActivity

B = int(os.environ.get("TB", "32"))
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = b200seg.MobileNetV2UNet(output_channels=10).to(dev).train()
eng = m._get_engine(); eng.precision = os.environ.get("B200SEG_TRAIN_PRECISION", "bf16"); eng.use_graphs = False
crit = b200seg.CrossEntropyLoss(); opt = torch.optim.Adam(m.parameters(), lr=1.5e-4)
x = torch.randn(B, 3, 256, 512, device=dev); y = torch.randint(0, 10, (B, 256, 512), device=dev)
def step():
    opt.zero_grad(); loss = crit(m(x), y); loss.backward(); opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
rows = [(e.key, e.count, e.device_time_total) for e in prof.key_averages() if e.device_time_total > 0]
rows.sort(key=lambda r: -r[2])
tot = sum(r[2] for r in rows)
print(f"total device time {tot/1e3:.2f} ms over {sum(r[1] for r in rows)} launches")
for k, n, t in rows[:40]:
    print(f"{t/1e3:9.3f} ms {t/tot:6.1%} x{n:4d}  {k[:100]}")
# per-launch durations of the reduction kernels (which layers are far from the roofline)
if os.environ.get("TP_DETAIL"):
    evs = [e for e in prof.events() if e.device_time_total > 0 and any(k in e.name for k in os.environ["TP_DETAIL"].split(","))]
    for e in evs:
        print(f"  {e.device_time_total:8.1f} us  {e.name[:60]}")
