"""Drop-in for the ``nn.CrossEntropyLoss()`` the reference builds at main.py:99 and calls at
train.py:37: same ``__call__(outputs, targets)``, mean reduction, no class weights; forward and
d(loss)/d(logits) come from ONE fused kernel (csrc/softmax_ce.cu)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _FusedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target):
        need = logits.requires_grad
        loss, dl = ops.softmax_ce(logits.contiguous(), target.contiguous(), want_grad=need)
        ctx.dl = dl
        return loss

    @staticmethod
    def backward(ctx, gout):
        dl, ctx.dl = ctx.dl, None
        if (gout.is_cuda and gout.numel() == 1 and gout.dtype == torch.float32 and dl.numel() % 4 == 0
                and not torch.is_grad_enabled()):
            # loss.backward(): gout is ones -- scale in place on the device only if it is not (no 2 x 168 MB elementwise pass)
            ops.scale_unless_one(dl, gout)
            return dl, None
        return dl * gout, None


class CrossEntropyLoss(nn.Module):
    """Mean per-pixel softmax cross-entropy over NCHW f32 logits and int64 [B,H,W] targets."""

    def forward(self, outputs: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        if not outputs.is_cuda:
            raise RuntimeError("b200seg.CrossEntropyLoss runs on CUDA only (no CPU fallback)")
        if outputs.dim() != 4:
            raise ValueError("expected NCHW logits")
        return _FusedCE.apply(outputs.float(), targets)
