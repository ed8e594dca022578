"""Per-step device time of the bench forward (B=64) with fused vs unfused inverted-residual blocks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch, b200seg
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = b200seg.MobileNetV2UNet(output_channels=10).to(dev).bfloat16().eval()
eng = model._get_engine()
xs = [torch.randn(64, 3, 256, 512).bfloat16().to(dev) for _ in range(4)]
for impl in ("unfused", "fused", None, "unfused", None):
    eng.mbconv_impl = impl
    with torch.no_grad():
        for i in range(6):
            model(xs[i % 4])
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(13)]
        evs[0].record()
        for i in range(12):
            model(xs[i % 4]); evs[i + 1].record()
        torch.cuda.synchronize()
    ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(12)]
    print(impl, "captured:", sum(e["graph"] is not None for e in eng._graphs.values()), "/", len(eng._graphs),
          " ms/step:", " ".join(f"{t:.2f}" for t in ts), flush=True)
