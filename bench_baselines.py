"""Baselines reported NEXT TO the product numbers by bench.py -- never on the product path.

  cpu_leg()        the reference's CPU path (fp32, batch 1, eval forward) on the host cores: the REAL reference
                   (/root/reference through the SURVEY Appendix-D shim) when that tree is present, else its restatement
                   oracle/unet_oracle.py (`kind` says which ran; the GPU box has no /root/reference).
  gpu_eager_leg()  the incumbent: the same model executed by stock torch eager (ATen -> cuDNN/cuBLAS) on the SAME B200
                   (main.py:13-21 picks `cuda` when it is there, so this is how the authors run it): config[1] forward in
                   bf16 channels_last and in fp32/TF32, config[2] training step under autocast(bf16) with torch.optim.Adam.
"""
import os
import statistics
import sys
import time

import torch

H, W, NCLS = 256, 512, 10
REF_ROOT = "/root/reference"


def load_reference_model():
    """MobileNetV2UNet(10) of the real reference (random init: unet.py:12 would download ImageNet weights), or None."""
    if not os.path.isdir(os.path.join(REF_ROOT, "src")):
        return None
    try:
        import torchvision.models as tvm
        orig = tvm.mobilenet_v2
        tvm.mobilenet_v2 = lambda *a, weights=None, **k: orig(*a, weights=None, **k)
        # by file path: `src.unet` on sys.path is this repo's own call-site shim (team02-objectdetection_b200/src/unet.py)
        import importlib.util
        spec = importlib.util.spec_from_file_location("_reference_src_unet", os.path.join(REF_ROOT, "src", "unet.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        torch.manual_seed(0)
        return mod.MobileNetV2UNet(output_channels=NCLS)
    except Exception:                                    # noqa: BLE001 -- torchvision missing, import error: use the port
        return None


def cpu_leg(seconds=12.0, warmup=2, fixed_iters=None):
    from oracle import unet_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    x = O.synth_input(1, H, W, seed=0)
    ref = load_reference_model()
    if ref is not None:
        ref.eval()
        fwd, kind, what = (lambda: ref(x)), "reference", "the reference's own src/unet.py MobileNetV2UNet (Appendix-D shim)"
    else:
        sd = O.synth_state_dict(O.mbv2unet_param_shapes(NCLS), seed=0)
        fwd, kind, what = (lambda: O.mobilenetv2_unet_forward(sd, x)), "port", "oracle port of src/unet.py"
    times = []
    with torch.no_grad():
        for _ in range(warmup):
            fwd()
        t_end = time.perf_counter() + seconds
        while (fixed_iters is None and time.perf_counter() < t_end) or (fixed_iters is not None and len(times) < fixed_iters):
            t0 = time.perf_counter()
            fwd()
            times.append(time.perf_counter() - t0)
    tot = sum(times)
    return dict(value=len(times) / tot, unit="images/s", cores=torch.get_num_threads(), kind=kind,
                sample=f"{len(times)} x (batch 1, 3x{H}x{W}, fp32 eval forward) {what} on PyTorch-CPU; "
                       f"best {min(times) * 1e3:.1f} ms median {statistics.median(times) * 1e3:.1f} ms",
                ms_per_image=tot / len(times) * 1e3)


def _time_cuda(fn, warmup, steps):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def gpu_eager_leg(dev, b_infer=64, b_train=32, warmup=3, steps=10):
    """Numbers only (images/s); every model here is stock torch.  Returns {} if anything is missing."""
    from oracle import unet_oracle as O
    out = {}
    ref = load_reference_model()
    out["eager_kind"] = "reference" if ref is not None else "port"
    g = torch.Generator().manual_seed(0)
    try:
        if ref is not None:
            def fwd_factory(dtype, cl):
                m = load_reference_model().to(dev).to(dtype).eval()
                if cl:
                    m = m.to(memory_format=torch.channels_last)
                return m
        else:
            sd = {k: v.to(dev) for k, v in O.synth_state_dict(O.mbv2unet_param_shapes(NCLS), seed=0).items()}

            def fwd_factory(dtype, cl):
                s2 = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
                if cl:
                    s2 = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in s2.items()}
                return lambda x: O.mobilenetv2_unet_forward(s2, x)
        # config[1]: eval forward, batch 64
        for tag, dtype, cl in (("eager_infer_bf16_cl_img_s", torch.bfloat16, True), ("eager_infer_fp32_tf32_img_s", torch.float32, False)):
            f = fwd_factory(dtype, cl)
            x = torch.randn(b_infer, 3, H, W, generator=g).to(dev).to(dtype)
            if cl:
                x = x.contiguous(memory_format=torch.channels_last)
            with torch.no_grad():
                ms = _time_cuda(lambda: f(x), warmup, steps)
            out[tag] = b_infer / (ms * 1e-3)
            del f, x
        # config[2]: training step, autocast bf16, torch.optim.Adam(lr=1.5e-4) (train.py:35-39, main.py:99-100)
        x = torch.randn(b_train, 3, H, W, generator=g).to(dev)
        y = torch.randint(0, NCLS, (b_train, H, W), generator=g).to(dev)
        crit = torch.nn.CrossEntropyLoss()
        if ref is not None:
            net = load_reference_model().to(dev).train()
            params = [p for p in net.parameters()]
            fwd = net
        else:
            pt = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k else v.clone()) for k, v in sd.items()}
            params = [v for v in pt.values() if v.requires_grad]

            def fwd(inp):
                return O.mobilenetv2_unet_forward(pt, inp, training=True, upd=O.BNState())
        opt = torch.optim.Adam(params, lr=1.5e-4)

        def step():
            opt.zero_grad()
            with torch.autocast("cuda", dtype=torch.bfloat16):
                o = fwd(x)
            loss = crit(o.float(), y)
            loss.backward()
            opt.step()
        ms = _time_cuda(step, warmup, steps)
        out["eager_train_autocast_bf16_img_s"] = b_train / (ms * 1e-3)
    except Exception as e:                               # noqa: BLE001 -- a baseline must never break the product bench
        out["eager_error"] = f"{type(e).__name__}: {str(e)[:90]}"
    finally:
        torch.cuda.empty_cache()
    return out
