"""Batch-sharded data parallelism for the drop-in models: one process per GPU, ``torch.distributed``
(NCCL over NVLink/NVSwitch; gloo for the CPU tests) as plumbing.

The reference is single-device (``main.py:13-21``); the parallel contract added here:
  * every rank holds a full replica and takes its slice of the batch; BatchNorm statistics stay per replica
    (the reference has no SyncBN);
  * the 194 used gradient tensors (26.2 MB fp32) live in ONE flat arena laid out in the order backward produces them
    (outc, up4 ... up1, then the encoder from features.18 down to the stem) and cut into buckets.  The backward
    kernels' partial results are turned into the arena (parameter layout, pre-scaled by 1/world) by one
    gradient-finalize launch per bucket, and the bucket's all-reduce is enqueued right behind it on a side stream --
    while the backward kernels of the earlier layers keep running on the main stream.  Under CUDA-graph replay the
    collectives are captured inside the backward graph (fork/join with events), so the overlap survives replay;
  * ``p.grad`` of every used parameter IS its arena view (no per-tensor copies after the all-reduce);
  * ``backbone.classifier`` is never on the forward path (SURVEY finding 5), gets no gradient and is excluded --
    a reducer that waited for it would hang;
  * parameters and buffers are broadcast from rank 0 once at start.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist


def used_parameters(model) -> List[torch.nn.Parameter]:
    """Parameters that receive gradients, in the order backward produces them (reverse schedule order)."""
    eng = model._get_engine() if hasattr(model, "_get_engine") else None
    if eng is None:
        return [p for p in model.parameters() if p.requires_grad]
    out, seen = [], set()
    for s in reversed(eng.steps):
        for mod in (s.bn, s.conv):
            if mod is None:
                continue
            for p in mod.parameters():
                if id(p) not in seen:
                    seen.add(id(p))
                    out.append(p)
    return out


class GradArena:
    """Flat fp32 gradient arena in backward order, cut into buckets of ~``bucket_bytes``.

    ``views[id(p)]`` is the parameter-shaped slice that becomes ``p.grad``.  ``reduce_bucket(i)`` starts the averaging
    all-reduce of bucket i (the producer has already scaled by 1/world, so the collective is a SUM: identical on NCCL and
    gloo); on CUDA it runs on ``self.side`` behind everything enqueued so far on the current stream.  ``join()`` makes the
    current stream (or the host, for gloo) wait for every collective started since the last join."""

    def __init__(self, params: List[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                 bucket_bytes: int = 8 << 20):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = list(params)
        dev = self.params[0].device
        self.offsets: Dict[int, int] = {}
        self.buckets: List[dict] = []          # {"params": [...], "lo": first element, "hi": end element}
        off, cur, lo = 0, [], 0
        for p in self.params:
            self.offsets[id(p)] = off
            cur.append(p)
            off += (p.numel() + 3) // 4 * 4    # every view starts on a 16-byte boundary
            if (off - lo) * 4 >= bucket_bytes:
                self.buckets.append(dict(params=cur, lo=lo, hi=off))
                cur, lo = [], off
        if cur:
            self.buckets.append(dict(params=cur, lo=lo, hi=off))
        self.flat = torch.zeros(max(off, 4), device=dev, dtype=torch.float32)
        self.views: Dict[int, torch.Tensor] = {
            id(p): self.flat[self.offsets[id(p)]:self.offsets[id(p)] + p.numel()].view(p.shape) for p in self.params}
        self.bucket_of: Dict[int, int] = {id(p): bi for bi, b in enumerate(self.buckets) for p in b["params"]}
        self.side = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self._works: list = []
        self._forked = False

    def bucket_flat(self, bi: int) -> torch.Tensor:
        b = self.buckets[bi]
        return self.flat[b["lo"]:b["hi"]]

    def reduce_bucket(self, bi: int) -> None:
        if self.world <= 1:
            return
        t = self.bucket_flat(bi)
        if self.side is None:                                   # gloo / CPU tensors
            self._works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            return
        self.side.wait_stream(torch.cuda.current_stream())      # fork: the bucket's finalize kernel is enqueued already
        with torch.cuda.stream(self.side):
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        self._forked = True

    def join(self) -> None:
        for w in self._works:
            w.wait()
        self._works = []
        if self._forked:
            torch.cuda.current_stream().wait_stream(self.side)
            self._forked = False


def broadcast_model(model, src: int = 0, group=None) -> None:
    """Rank `src`'s parameters and buffers (BN running stats, num_batches_tracked) to every replica."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    seen = set()
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            if t.data_ptr() in seen:      # downK.N.* aliases backbone.features.N.*
                continue
            seen.add(t.data_ptr())
            dist.broadcast(t, src=src, group=group)


def attach(model, group=None, bucket_bytes: int = 8 << 20) -> None:
    """Make ``loss.backward()`` of `model` produce rank-averaged gradients (bucketed all-reduce overlapped with
    backward, also under CUDA-graph replay).  Call after ``init_process_group`` and ``broadcast_model``."""
    eng = model._get_engine()
    eng.dp = dict(group=group, bucket_bytes=bucket_bytes)
    eng._train_ws = None                  # the gradient arena is rebuilt with the new bucket plan
    eng._graphs.clear()


def detach(model) -> None:
    """Back to single-replica gradients; releases every captured step graph.  NCCL refuses to tear a communicator down
    while CUDA graphs that captured its collectives are alive, so call this (or drop the model) BEFORE
    ``dist.destroy_process_group()``."""
    import gc
    eng = model._get_engine()
    eng.dp = None
    eng._train_ws = None
    eng._graphs.clear()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
