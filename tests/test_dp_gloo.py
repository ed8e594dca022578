"""T4 host logic on CPU: the bucketed gradient reducer and the parameter broadcast,
world_size 2 over gloo.  (The model itself needs CUDA; here the reducer is driven with synthetic gradients in
the exact order backward produces them.)"""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import b200seg
from b200seg import dp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                       # different replicas before the broadcast
        m = b200seg.UNet(output_channels=10, base_filters=16)
        dp.broadcast_model(m, src=0)
        w0 = m.inc.conv.conv[0].weight.detach().clone()
        # build the schedule without touching CUDA: the Engine constructor is pure Python
        from b200seg import engine
        m.__dict__["_engine"] = engine.Engine(m, "unet")
        params = dp.used_parameters(m)
        red = dp.GradBucketReducer(params, bucket_bytes=64 << 10)
        grads = {id(p): torch.full_like(p, float(rank + 1)) * (i + 1) for i, p in enumerate(params)}
        red.reset()
        for p in params:                                    # backward order
            red.add(p, grads[id(p)])
        avg = red.finish()
        ok = all(torch.allclose(avg[id(p)], torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(params))
        q.put((rank, ok, float(w0.sum()), len(red.buckets), len(params)))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_and_broadcast_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res                      # every gradient == mean over ranks
    assert res[0][2] == res[1][2]                           # broadcast made the replicas identical
    assert res[0][3] > 1                                    # more than one bucket
    assert res[0][4] == 62                                  # 16 convs (weight+bias) + 15 BNs (weight+bias)


def test_unused_classifier_is_excluded_from_the_reducer():
    m = b200seg.MobileNetV2UNet(output_channels=10)
    from b200seg import engine
    m.__dict__["_engine"] = engine.Engine(m, "mbv2unet")
    params = dp.used_parameters(m)
    assert len(params) == 194                               # SURVEY: 196 tensors, 194 ever get gradients
    ids = {id(p) for p in params}
    assert id(m.backbone.classifier[1].weight) not in ids and id(m.backbone.classifier[1].bias) not in ids
    assert sum(p.numel() for p in params) == 6549786        # the 26.2 MB all-reduce payload
    # backward order: the decoder/outc gradients come first, the stem last
    names = {id(p): n for n, p in m.named_parameters()}
    order = [names[id(p)] for p in params]
    assert order[0].startswith("outc.") and order[-1].startswith("backbone.features.0.")
    red = dp.GradBucketReducer(params, bucket_bytes=8 << 20)
    assert 2 <= len(red.buckets) <= 6
