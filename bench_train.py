"""Training-step benchmark (BASELINE config[2]): MobileNetV2UNet fwd + pixel CE + bwd + Adam, batch 32 per GPU at
3x256x512, data parallel over the ranks.  Invoked as `python bench.py --workload train ...` (same launch contract
as the inference bench; see bench.py).

A step is exactly the reference's loop body (train.py:35-39): optimizer.zero_grad(); outputs = model(inputs);
loss = criterion(outputs, targets); loss.backward(); optimizer.step() -- with torch.optim.Adam(lr=1.5e-4)
(main.py:100) and, for N>1, the bucketed gradient all-reduce of b200seg.dp overlapped with backward.
"""
import json
import os
import sys

import torch

H, W, NCLS = 256, 512, 10


def run_train(args, dev, dist, world, rank, pk, clocks, emit):
    import b200seg
    from b200seg import dp, train_path
    global H, W
    unet = args.workload == "unet_train"
    if unet:
        H, W = 512, 1024
    B = args.batch or (4 if unet else 32)
    precision = os.environ.get("B200SEG_TRAIN_PRECISION", "bf16")     # bf16 activations + tcgen05 convs, fp32 master weights
    torch.manual_seed(0)
    model = (b200seg.UNet(output_channels=NCLS) if unet else b200seg.MobileNetV2UNet(output_channels=NCLS)).to(dev)
    eng = model._get_engine()
    eng.precision = precision
    if dist is not None:
        dp.broadcast_model(model)
        dp.attach(model)
    criterion = b200seg.CrossEntropyLoss()
    fused_adam = os.environ.get("B200SEG_BENCH_ADAM", "fused") == "fused"
    # main.py:100 -- same constructor call; b200seg.Adam is the one-launch multi-tensor drop-in for optim.Adam
    optimizer = (b200seg.Adam if fused_adam else torch.optim.Adam)(model.parameters(), lr=1.5e-4)
    model.train()
    g = torch.Generator(device="cpu").manual_seed(rank)
    nrot = 2
    xs = [torch.randn(B, 3, H, W, generator=g).to(dev) for _ in range(nrot)]
    ys = [torch.randint(0, NCLS, (B, H, W), generator=g).to(dev) for _ in range(nrot)]

    def step(x, y):
        optimizer.zero_grad()
        out = model(x)
        loss = criterion(out, y)
        loss.backward()
        optimizer.step()
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):        # (the nvidia-smi sampler was started by bench.py, seconds ago)
        step(xs[i % nrot], ys[i % nrot])
    barrier()
    clocks.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import time
    e0.record()
    th0 = time.perf_counter()
    for i in range(args.steps):
        loss = step(xs[i % nrot], ys[i % nrot])
    host_issue_ms = (time.perf_counter() - th0) * 1e3 / args.steps      # host time to ISSUE a step (no sync inside)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    last_loss = float(loss)

    # e2e: host (pinned) fp32 images + int64 labels in, loss.item() out, every step
    xh = [torch.randn(B, 3, H, W, generator=g).pin_memory() for _ in range(2)]
    yh = [torch.randint(0, NCLS, (B, H, W), generator=g).pin_memory() for _ in range(2)]

    # input feed: pinned host batches, non-blocking uploads on a copy stream, one batch ahead -- the
    # upload of batch i+1 overlaps step i although the loss is read back (a sync) every step as train.py:41-42 does
    copy_s, comp_s = torch.cuda.Stream(), torch.cuda.current_stream()
    xd = [torch.empty(B, 3, H, W, device=dev) for _ in range(2)]
    yd = [torch.empty(B, H, W, dtype=torch.int64, device=dev) for _ in range(2)]

    def upload(i):
        k = i & 1
        with torch.cuda.stream(copy_s):
            copy_s.wait_stream(comp_s)              # buffer k was last read by step i-2, already enqueued on comp_s
            xd[k].copy_(xh[k], non_blocking=True)
            yd[k].copy_(yh[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_s)
        return ev

    def e2e_steps(n):
        ev = upload(0)
        for i in range(n):
            comp_s.wait_event(ev)
            if i + 1 < n:
                ev_next = upload(i + 1)
            step(xd[i & 1], yd[i & 1]).item()       # train.py:41-42 reads the loss every step
            if i + 1 < n:
                ev = ev_next

    e2e_steps(3)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_steps(args.steps)
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)

    # per-layer forward/backward times (CUDA events on the launch stream), one extra step
    train_path.TRACE = []
    step(xs[0], ys[0])
    torch.cuda.synchronize()
    tr, train_path.TRACE = train_path.TRACE, None
    agg = {}
    for phase, name, a, b in tr:
        agg[(phase, name)] = agg.get((phase, name), 0.0) + a.elapsed_time(b)
    rows = sorted(agg.items(), key=lambda kv: -kv[1])
    tot = sum(v for _, v in rows)
    if args.breakdown and rank == 0:
        print(f"{'phase':4s} {'layer':34s} {'ms':>9s} {'share':>6s}", file=sys.stderr)
        for (phase, name), v in rows[:30]:
            print(f"{phase:4s} {name:34s} {v:9.3f} {v / tot:6.1%}", file=sys.stderr)
        print(f"sum fwd {sum(v for (p, _), v in rows if p == 'fwd'):.2f} ms, bwd {sum(v for (p, _), v in rows if p == 'bwd'):.2f} ms; "
              f"step {ms / args.steps:.2f} ms", file=sys.stderr)

    if dist is not None:
        t = torch.tensor([ms, ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    (top_phase, top_name), top_ms = rows[0]
    # algorithmic model of the whole step: 353 MB/img bf16 activations (706 MB fp32), 34.5 GFLOP/img
    bytes_img = (4772e6 if unet else 353e6) * (1 if precision == "bf16" else 2)
    step_s = ms / args.steps * 1e-3
    line = {"metric": ("UNet" if unet else "MobileNetV2UNet") + " training images/s", "value": B * world * args.steps / (ms * 1e-3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"{'UNet' if unet else 'MobileNetV2UNet'} training step (fwd + pixel CE + bwd + Adam), batch {B}/GPU, 3x{H}x{W}, {NCLS} classes "
                                   f"(BASELINE config[{3 if unet else 2}]); activations {precision}, fp32 master weights, {'b200seg.Adam (fused multi-tensor)' if fused_adam else 'torch.optim.Adam'}(lr=1.5e-4); "
                                   f"data parallel, per-replica BatchNorm, bucketed all-reduce overlapped with backward",
                       "global_batch": B * world, "l2": "~20 GB of activations per step >> 126 MB L2", "last_loss": last_loss,
                       "host_issue_ms_per_step": host_issue_ms},
            "e2e": {"value": B * world * args.steps / (ms_e2e * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": B * 3 * H * W * 4 + B * H * W * 8, "d2h_bytes_per_step": 4,
                    "api": "train.py:32-42 loop body: pinned fp32 images + int64 labels uploaded one batch ahead on a copy stream, step, loss.item() every step"},
            "roofline": {"kernel": f"{top_phase}:{top_name} (layer-level; dense weight gradients still run on the FP32 pipes)",
                         "bound": "hbm", "achieved": bytes_img * B / step_s / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                         "frac": bytes_img * B / step_s / 1e9 / pk["hbm"], "traffic": None, "peak_source": pk["src"],
                         "note": "whole-step algorithmic bytes / step time; per-layer shares in --breakdown",
                         "share_of_step": top_ms / tot},
            "gpu_launches": None, "clocks": clk}
    emit(line)                # bench.py's emitter: the ONE stdout line (bench.py runs as __main__, do not re-import it)
    if dist is not None:
        dist.destroy_process_group()
