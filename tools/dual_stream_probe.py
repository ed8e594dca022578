"""Experiment: does running the two halves of the batch as two parallel branches of the forward graph (two streams)
hide the tails / launch gaps of the many small kernels?  Prints ms per forward body for 1, 2 and 4 branches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch, b200seg
from b200seg import ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
B = 64
model = b200seg.MobileNetV2UNet(output_channels=10).to(dev).bfloat16().eval()
eng = model._get_engine()
x = torch.randn(B, 3, 256, 512).bfloat16().to(dev)
mode, sdt, dense_impl = "bf16", torch.bfloat16, "tc"
pk = eng._pack_eval(mode)
steps = eng._schedule(mode, dense_impl, 256, 512)
args = (pk, mode, sdt, dense_impl, torch.bfloat16, False)
with torch.no_grad():
    env0 = {"x": x}
    eng._run_step(steps[0], env0, *args)
    f0 = env0[steps[0].dst]
    for nb in (1, 2, 4):
        parts = list(f0.chunk(nb, 0))
        # warm-up eager (lazy init of kernels for these shapes)
        for p in parts:
            env = {steps[0].dst: p}
            for s in steps[1:-1]:
                eng._run_step(s, env, *args)
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream() for _ in range(nb)]
        g = torch.cuda.CUDAGraph()
        keep = []
        with torch.cuda.graph(g):
            cur = torch.cuda.current_stream()
            for st, p in zip(streams, parts):
                st.wait_stream(cur)
                with torch.cuda.stream(st):
                    env = {steps[0].dst: p}
                    for s in steps[1:-1]:
                        eng._run_step(s, env, *args)
                    keep.append(env)
            for st in streams:
                cur.wait_stream(st)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"branches={nb}: {e0.elapsed_time(e1) / 10:.3f} ms per forward body (B={B})", flush=True)
