"""Batch-sharded data parallelism for the drop-in models: one process per GPU, ``torch.distributed``
(NCCL over NVLink/NVSwitch; gloo for the CPU tests) as plumbing.

The reference is single-device (``main.py:13-21``); the parallel contract added here:
  * every rank holds a full replica and takes its slice of the batch; BatchNorm statistics stay per replica
    (the reference has no SyncBN);
  * after backward the 194 used gradient tensors (26.2 MB fp32) are averaged with a bucketed all-reduce.
    Buckets are filled in the order backward produces gradients (outc, up4 ... up1, then the encoder from
    features.18 down to the stem), and each bucket's all-reduce is launched asynchronously as soon as it is
    full, i.e. while the backward kernels of earlier layers are still running;
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


class GradBucketReducer:
    """Flat-bucket gradient averaging.  ``add(param, grad)`` copies (pre-scaled by 1/world) into the bucket and,
    when the bucket is complete, starts its all-reduce; ``finish()`` waits and returns {id(param): averaged view}."""

    def __init__(self, params: List[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                 bucket_bytes: int = 8 << 20):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = list(params)
        self.buckets: List[dict] = []
        cur, size = [], 0
        for p in self.params:
            cur.append(p)
            size += p.numel() * 4
            if size >= bucket_bytes:
                self._close(cur)
                cur, size = [], 0
        if cur:
            self._close(cur)
        self.where: Dict[int, tuple] = {}
        for bi, b in enumerate(self.buckets):
            off = 0
            for p in b["params"]:
                self.where[id(p)] = (bi, off, p.numel(), tuple(p.shape))
                off += p.numel()
        self.reset()

    def _close(self, plist):
        n = sum(p.numel() for p in plist)
        dev = plist[0].device
        self.buckets.append(dict(params=list(plist), flat=torch.zeros(n, device=dev, dtype=torch.float32), pending=0,
                                 work=None))

    def reset(self):
        for b in self.buckets:
            b["pending"] = len(b["params"])
            b["work"] = None

    def add(self, p: torch.nn.Parameter, grad: torch.Tensor) -> None:
        bi, off, n, _ = self.where[id(p)]
        b = self.buckets[bi]
        dst = b["flat"][off:off + n]
        torch.mul(grad.reshape(-1).to(torch.float32), 1.0 / self.world, out=dst)
        b["pending"] -= 1
        if b["pending"] == 0 and self.world > 1:
            b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self) -> Dict[int, torch.Tensor]:
        out = {}
        for b in self.buckets:
            if b["pending"] != 0:
                raise RuntimeError("gradient bucket incomplete: a parameter on the path produced no gradient")
            if b["work"] is not None:
                b["work"].wait()
        for pid, (bi, off, n, shape) in self.where.items():
            out[pid] = self.buckets[bi]["flat"][off:off + n].view(shape)
        return out


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


def attach(model, group=None, bucket_bytes: int = 8 << 20) -> GradBucketReducer:
    """Make ``loss.backward()`` of `model` produce rank-averaged gradients (all-reduce overlapped with backward)."""
    eng = model._get_engine()
    red = GradBucketReducer(used_parameters(model), group, bucket_bytes)
    eng.reducer = red
    return red
