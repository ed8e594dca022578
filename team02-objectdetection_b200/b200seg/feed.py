"""Input feed for the training loop: what ``train.py:32-33`` does with two blocking pageable copies per
batch --

    for inputs, targets in train_bar:
        inputs = inputs.to(device)
        targets = targets.to(device)

-- as a double-buffered, pinned, asynchronous upload one batch ahead of the step, plus the datasets' label remap
(``BDD100KDataset.py:23-35,66-69``: ``mapped_mask[mask == src] = target`` per class, then ``.long()``) on the GPU so that
only uint8 labels cross PCIe (1 byte per pixel instead of 8).

    feeder = b200seg.DeviceFeeder(train_loader, device)            # drop-in for the loop header
    for inputs, targets in feeder:                                 # already on the device, upload of the next batch in flight
        ...

    lut = b200seg.class_map_lut({0: 1, 13: 2, 6: 3, ...})          # the dataset's class_map
    targets = b200seg.remap_labels(mask_u8_on_gpu, lut)            # int64 [B,H,W], equal to the reference's loop + .long()
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, Optional, Tuple

import torch

from ._cabi import check, lib, ptr


def class_map_lut(class_map: Dict[int, int], default: int = 0, device=None) -> torch.Tensor:
    """256-entry uint8 table equal to the reference's loop ``mapped = zeros; for s, t in class_map.items(): mapped[mask == s] = t``
    (later entries win, like the loop; unmapped sources -> ``default`` = background)."""
    lut = torch.full((256,), int(default), dtype=torch.uint8)
    for src, dst in class_map.items():
        if not 0 <= int(src) <= 255 or not 0 <= int(dst) <= 255:
            raise ValueError(f"class_map entry {src}: {dst} does not fit uint8 labels")
        lut[int(src)] = int(dst)
    return lut.to(device) if device is not None else lut


def remap_labels(mask: torch.Tensor, lut: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``lut[mask]`` as int64 on the GPU: uint8 label image(s) -> CrossEntropyLoss targets (csrc/preprocess.cu)."""
    if not mask.is_cuda:
        raise RuntimeError("remap_labels runs on CUDA tensors (upload the uint8 mask first; no CPU fallback)")
    if mask.dtype != torch.uint8:
        raise TypeError("remap_labels expects uint8 labels")
    if lut.dtype != torch.uint8 or lut.numel() != 256:
        raise ValueError("lut must be a 256-entry uint8 table (class_map_lut)")
    lut = lut.to(mask.device).contiguous()
    mask = mask.contiguous()
    if out is None:
        out = torch.empty(mask.shape, dtype=torch.int64, device=mask.device)
    elif out.shape != mask.shape or out.dtype != torch.int64 or not out.is_contiguous() or out.device != mask.device:
        raise ValueError("remap_labels: bad output tensor")
    if mask.numel():
        with torch.cuda.device(mask.device):
            check(lib.b200seg_remap_labels(ptr(mask), ptr(out), ptr(lut), mask.numel(), torch.cuda.current_stream().cuda_stream),
                  "remap_labels")
    return out


class DeviceFeeder:
    """Iterates a loader of (inputs, targets) host batches and yields them on ``device``, uploading batch i+1 on a copy
    stream while the caller's step on batch i runs.  Host batches are staged in two pinned buffers per tensor (allocated
    for the largest batch seen); device buffers are double-buffered too, so a yielded pair stays valid until the batch
    after the next one is requested -- exactly the lifetime ``train.py``'s loop body needs.

    ``lut``: optional class_map table; targets that arrive as uint8 are remapped on the device (``remap_labels``) and
    yielded as int64.  Float inputs keep their dtype; int64 targets are uploaded as they are."""

    def __init__(self, loader: Iterable, device, lut: Optional[torch.Tensor] = None):
        self.loader, self.device = loader, torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceFeeder feeds a CUDA device")
        self.lut = lut.to(self.device) if lut is not None else None
        self.copy_stream = torch.cuda.Stream(self.device)
        self._pin = [[None, None], [None, None]]          # [slot][inputs | targets]
        self._dev = [[None, None], [None, None]]
        self._consumed = [None, None]                     # event: the step that used slot k has been enqueued

    def __len__(self):
        return len(self.loader)

    @staticmethod
    def _fit(buf, like: torch.Tensor, **kw):
        if buf is None or buf.dtype != like.dtype or buf.numel() < like.numel():
            buf = torch.empty(like.numel(), dtype=like.dtype, **kw)
        return buf

    def _upload(self, k: int, batch) -> Tuple[torch.Tensor, torch.Tensor, torch.cuda.Event]:
        out = []
        with torch.cuda.stream(self.copy_stream):
            if self._consumed[k] is not None:
                self.copy_stream.wait_event(self._consumed[k])       # slot k's previous batch is no longer being read
            for t_i, t in enumerate(batch[:2]):
                t = t.contiguous()
                self._dev[k][t_i] = self._fit(self._dev[k][t_i], t, device=self.device)
                d = self._dev[k][t_i][:t.numel()].view(t.shape)
                if t.is_pinned():                                    # DataLoader(pin_memory=True): upload straight from it
                    h = t
                else:
                    self._pin[k][t_i] = self._fit(self._pin[k][t_i], t, pin_memory=True)
                    h = self._pin[k][t_i][:t.numel()].view(t.shape)
                    h.copy_(t)                                       # pageable -> pinned on the host (what .to() hides inside)
                d.copy_(h, non_blocking=True)
                out.append(d)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return out[0], out[1], ev

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        it = iter(self.loader)
        try:
            nxt = self._upload(0, next(it))
        except StopIteration:
            return
        k = 0
        while nxt is not None:
            x, y, ev = nxt
            try:
                nxt = self._upload(k ^ 1, next(it))                  # in flight while the caller works on (x, y)
            except StopIteration:
                nxt = None
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            if self.lut is not None and y.dtype == torch.uint8:
                y = remap_labels(y, self.lut)
            yield x, y
            done = torch.cuda.Event()
            done.record(torch.cuda.current_stream(self.device))      # everything the caller enqueued on (x, y)
            self._consumed[k] = done
            k ^= 1
