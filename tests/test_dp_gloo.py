"""T4 host logic on CPU: the flat bucketed gradient arena and the parameter broadcast,
world_size 2 over gloo.  (The model itself needs CUDA; here the arena is filled with synthetic, 1/world-scaled
gradients bucket by bucket in the exact order backward finalizes them.)"""
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
        arena = dp.GradArena(params, bucket_bytes=64 << 10)
        index = {id(p): i for i, p in enumerate(params)}
        for bi, b in enumerate(arena.buckets):              # backward order: what grad_finalize does per bucket on the GPU
            for p in b["params"]:
                arena.views[id(p)].fill_(float(rank + 1) * (index[id(p)] + 1) / world)
            arena.reduce_bucket(bi)                         # async: later buckets are filled while this one is in flight
        arena.join()
        ok = all(torch.allclose(arena.views[id(p)], torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(params))
        # views tile the flat buffer in order, 16-byte aligned, buckets are contiguous and cover everything
        ok = ok and all(arena.offsets[id(p)] % 4 == 0 for p in params)
        ok = ok and arena.buckets[0]["lo"] == 0 and all(a["hi"] == b["lo"] for a, b in zip(arena.buckets, arena.buckets[1:]))
        ok = ok and arena.buckets[-1]["hi"] == arena.flat.numel()
        q.put((rank, ok, float(w0.sum()), len(arena.buckets), len(params)))
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


def test_unused_classifier_is_excluded_from_the_arena():
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
    arena = dp.GradArena(params, bucket_bytes=8 << 20)
    assert 2 <= len(arena.buckets) <= 6
    assert all(arena.views[id(p)].shape == p.shape for p in params)
