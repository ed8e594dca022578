"""cProfile of the HOST side of the graphed training step (what train_host_issue_ms is made of)."""
import os, sys, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
import b200seg
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = b200seg.MobileNetV2UNet(output_channels=10).to(dev).train()
eng = m._get_engine(); eng.precision = "bf16"
crit = b200seg.CrossEntropyLoss(); opt = b200seg.Adam(m.parameters(), lr=1.5e-4)
x = torch.randn(32, 3, 256, 512, device=dev); y = torch.randint(0, 10, (32, 256, 512), device=dev)
def step():
    opt.zero_grad(); loss = crit(m(x), y); loss.backward(); opt.step(); return loss
for _ in range(6): step()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(20): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host issue {(t1 - t0) / 20 * 1e3:.2f} ms/step, wall {(t2 - t0) / 20 * 1e3:.2f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(35); print(s.getvalue()[:6000])
