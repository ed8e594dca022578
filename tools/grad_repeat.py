import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "team02-objectdetection_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch, b200seg
from oracle import unet_oracle as O
from util import expand_aliases, fixture_sd
from test_gpu_train import _oracle_step
sd = fixture_sd()
x, t = O.synth_input(2, 64, 64, seed=1), O.synth_target(2, 64, 64, seed=1)
_, _, p32, _ = _oracle_step(sd, x, t)
_, _, p64, _ = _oracle_step(sd, x, t, dt=torch.float64)
prev = None
for rep in range(4):
    m = b200seg.MobileNetV2UNet(output_channels=10); m.load_state_dict(expand_aliases(sd), strict=True); m = m.to("cuda").train()
    loss = b200seg.CrossEntropyLoss()(m(x.cuda()), t.cuda()); loss.backward()
    g = {n: p.grad.detach().cpu().double() for n, p in m.named_parameters() if p.grad is not None}
    errs = {}
    for n, gg in g.items():
        ref = p64[n].grad; sc = float(ref.abs().max())
        if sc < 1e-5: continue
        errs[n] = float((gg - ref).abs().max()) / sc
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    same = None if prev is None else max(float((g[n] - prev[n]).abs().max() / prev[n].abs().max().clamp_min(1e-30)) for n in g)
    print(rep, "loss", float(loss.detach()), "top", [(n, round(e, 4)) for n, e in top], "max rel change vs previous run", same)
    prev = g
