"""Cycles per tcgen05.mma (kind::f16, K = 16) as a function of M, N and CTAs per SM (DESIGN.md finding 8)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200"))
import torch
from b200seg._cabi import ptr
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _toolslib import check, lib
ITERS = 4096
sms = torch.cuda.get_device_properties(0).multi_processor_count
print(f"{'M':>4s} {'N':>4s} {'CTAs/SM':>8s} {'cycles/MMA per CTA':>20s} {'cycles/MMA per SM':>18s} {'MAC/clk/SM':>11s}")
for M in (128, 64):
    for N in (16, 32, 48, 64, 96, 128, 192, 240, 256):
        for ctas in (1, 2):
            out = torch.zeros(sms * ctas, dtype=torch.int64, device="cuda")
            for _ in range(2):
                check(lib.b200seg_probe_mma(M, N, ITERS, ctas, ptr(out), torch.cuda.current_stream().cuda_stream), "probe_mma")
            torch.cuda.synchronize()
            c = float(out.double().mean()) / ITERS
            per_sm = c / ctas
            print(f"{M:4d} {N:4d} {ctas:8d} {c:20.1f} {per_sm:18.1f} {M * N * 16 / per_sm:11.0f}", flush=True)
