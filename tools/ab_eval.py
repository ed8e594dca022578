"""A/B timing of the eval forward (BASELINE config[1]: B=64, 3x256x512, bf16) under different engine settings, same process,
interleaved.  usage: python tools/ab_eval.py "tail_impl=unfused" "tail_impl=None" ...   (each arg: k=v[,k=v] on the Engine)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200")); sys.path.insert(0, ROOT)
import torch
import b200seg

B, H, W = int(os.environ.get("AB_B", 64)), 256, 512
variants = sys.argv[1:] or ["tail_impl=None"]
torch.manual_seed(0)
m = b200seg.MobileNetV2UNet(output_channels=10).cuda().bfloat16().eval()
eng = m._get_engine()
xs = [torch.randn(B, 3, H, W).bfloat16().cuda() for _ in range(4)]
mask = os.environ.get("AB_MASK") == "1"


def setup(v):
    for kv in v.split(","):
        k, val = kv.split("=")
        setattr(eng, k, None if val == "None" else (int(val, 0) if val.lstrip("-").replace("x", "").isalnum() and val[0].isdigit() else val))


def run(n):
    for i in range(n):
        y = m.predict_mask(xs[i % 4]) if mask else m(xs[i % 4])
    return y


res = {v: [] for v in variants}
with torch.no_grad():
    for rep in range(3):
        for v in variants:
            setup(v)
            run(6)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(40); e1.record(); torch.cuda.synchronize()
            res[v].append(e0.elapsed_time(e1) / 40)
for v, t in res.items():
    print(f"{v:40s} ms/step {min(t):.4f} (runs {' '.join('%.4f' % x for x in t)})  {B / min(t) * 1e3:.0f} img/s", flush=True)
