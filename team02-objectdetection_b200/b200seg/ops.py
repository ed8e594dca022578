"""Tensor-level wrappers over the C ABI: one function per ATen dispatch of the reference hot path
(SURVEY.md section 2.1).  Each validates device/dtype/layout, passes raw pointers and the current
CUDA stream, and returns the output tensor.  Activations are NHWC-contiguous torch tensors of shape
[B, H, W, C]; nothing here ever falls back to torch arithmetic.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _cabi
from ._cabi import ACT_NONE, ACT_RELU, ACT_RELU6, BF16, F32, check, lib, ptr  # noqa: F401

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"b200seg supports float32 and bfloat16 activations, got {t.dtype}") from None


def _cuda(*ts):
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("b200seg ops need CUDA tensors (no CPU fallback)")
        if not t.is_contiguous():
            raise RuntimeError("b200seg ops need contiguous tensors")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def conv3x3_smallcin(x_nchw, w, b, stride: int, act: int, out_dtype, out=None):
    """x NCHW [B,Cin<=4,H,W]; w f32 [3,3,Cin,Cout]; -> NHWC [B,Ho,Wo,Cout]."""
    _cuda(x_nchw, w, b)
    B, Cin, H, W = x_nchw.shape
    Cout = w.shape[-1]
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    if out is None:
        out = torch.empty((B, Ho, Wo, Cout), device=x_nchw.device, dtype=out_dtype)
    check(lib.b200seg_conv3x3_smallcin(ptr(x_nchw), _dt(x_nchw), ptr(w), ptr(b), ptr(out), _dt(out), B, Cin, H, W,
                                       Cout, stride, act, _stream()), "conv3x3_smallcin")
    return out


def dwconv3x3(x, w9c, b, stride: int, act: int, out=None):
    """x NHWC; w9c f32 [9,C]."""
    _cuda(x, w9c, b)
    B, H, W, Cc = x.shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    if out is None:
        out = torch.empty((B, Ho, Wo, Cc), device=x.device, dtype=x.dtype)
    check(lib.b200seg_dwconv3x3(ptr(x), ptr(w9c), ptr(b), ptr(out), _dt(x), B, H, W, Cc, stride, act, _stream()),
          "dwconv3x3")
    return out


def pack_dw_diag(w9c: torch.Tensor) -> torch.Tensor:
    """f32 [9, C] depthwise taps -> block-diagonal bf16 [C, 9, 64] for dwconv3x3_tc."""
    C = w9c.shape[1]
    out = torch.zeros(C, 9, 64, device=w9c.device, dtype=torch.bfloat16)
    idx = torch.arange(C, device=w9c.device)
    out[idx, :, idx % 64] = w9c.t().to(torch.bfloat16)
    return out.contiguous()


def dwconv3x3_tc(x, wdiag, b, stride: int, act: int, out=None, flags: int = 0):
    """Depthwise 3x3 on the tensor cores. x NHWC bf16; wdiag from pack_dw_diag."""
    _cuda(x, wdiag, b)
    if x.dtype != torch.bfloat16 or wdiag.dtype != torch.bfloat16:
        raise TypeError("dwconv3x3_tc is bf16-only")
    B, H, W, Cc = x.shape
    if tuple(wdiag.shape) != (Cc, 9, 64):
        raise ValueError(f"dwconv3x3_tc: wdiag {tuple(wdiag.shape)} != ({Cc}, 9, 64)")
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    if out is None:
        out = torch.empty((B, Ho, Wo, Cc), device=x.device, dtype=torch.bfloat16)
    check(lib.b200seg_dwconv3x3_tc(ptr(x), ptr(wdiag), ptr(b), ptr(out), B, H, W, Cc, stride, act, flags, _stream()),
          "dwconv3x3_tc")
    return out


def conv_tc(x, w, b, taps: int, act: int, res=None, out=None, flags: int = 0):
    """Tensor-core conv. x NHWC bf16 [B,H,W,Cin]; w bf16 [Cout, taps*Cin]; b f32 [Cout]."""
    _cuda(x, w, b, res)
    if x.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise TypeError("conv_tc is bf16-only")
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    if w.shape[1] != taps * Cin:
        raise ValueError(f"conv_tc: weight {tuple(w.shape)} does not match taps={taps} Cin={Cin}")
    if out is None:
        out = torch.empty((B, H, W, Cout), device=x.device, dtype=torch.bfloat16)
    check(lib.b200seg_conv_tc(ptr(x), ptr(w), ptr(b), ptr(res), ptr(out), B, H, W, Cin, Cout, taps, act, flags,
                              _stream()), "conv_tc")
    return out


def conv_simt(x, w, b, taps: int, act: int, res=None, out=None):
    """FP32-pipe conv. x NHWC (f32|bf16); w f32 [Cout, taps*Cin]."""
    _cuda(x, w, b, res)
    if w.dtype != torch.float32:
        raise TypeError("conv_simt weights are f32")
    B, H, W, Cin = x.shape
    Cout = w.shape[0]
    if w.shape[1] != taps * Cin:
        raise ValueError(f"conv_simt: weight {tuple(w.shape)} does not match taps={taps} Cin={Cin}")
    if out is None:
        out = torch.empty((B, H, W, Cout), device=x.device, dtype=x.dtype)
    check(lib.b200seg_conv_simt(ptr(x), ptr(w), ptr(b), ptr(res), ptr(out), _dt(x), B, H, W, Cin, Cout, taps, act,
                                _stream()), "conv_simt")
    return out


def upsample2x_concat(skip, x, out=None):
    """cat([skip, bilinear_x2(x, align_corners=False)], channel) in NHWC."""
    _cuda(skip, x)
    B, h, w, Cu = x.shape
    Cs = skip.shape[3]
    if skip.shape[:3] != (B, 2 * h, 2 * w) or skip.dtype != x.dtype:
        raise ValueError(f"upsample2x_concat: skip {tuple(skip.shape)} vs x {tuple(x.shape)}")
    if out is None:
        out = torch.empty((B, 2 * h, 2 * w, Cs + Cu), device=x.device, dtype=x.dtype)
    check(lib.b200seg_upsample2x_concat(ptr(skip), ptr(x), ptr(out), _dt(x), B, h, w, Cs, Cu, _stream()),
          "upsample2x_concat")
    return out


def upsample2x_ac_nchw(logits, C: int, out_dtype, out=None):
    """NHWC logits [B,h,w,ldc] -> NCHW [B,C,2h,2w], bilinear align_corners=True."""
    _cuda(logits)
    B, h, w, ldc = logits.shape
    if out is None:
        out = torch.empty((B, C, 2 * h, 2 * w), device=logits.device, dtype=out_dtype)
    check(lib.b200seg_upsample2x_ac_nchw(ptr(logits), _dt(logits), ldc, ptr(out), _dt(out), B, h, w, C, _stream()),
          "upsample2x_ac_nchw")
    return out


def upsample2x_ac_argmax(logits, C: int, out=None):
    _cuda(logits)
    B, h, w, ldc = logits.shape
    if out is None:
        out = torch.empty((B, 2 * h, 2 * w), device=logits.device, dtype=torch.uint8)
    check(lib.b200seg_upsample2x_ac_argmax(ptr(logits), _dt(logits), ldc, ptr(out), B, h, w, C, _stream()),
          "upsample2x_ac_argmax")
    return out


def nhwc_to_nchw(x, C: int, out_dtype, out=None):
    _cuda(x)
    B, H, W, ldc = x.shape
    if out is None:
        out = torch.empty((B, C, H, W), device=x.device, dtype=out_dtype)
    check(lib.b200seg_nhwc_to_nchw(ptr(x), _dt(x), ldc, ptr(out), _dt(out), B, H, W, C, _stream()), "nhwc_to_nchw")
    return out


def maxpool2x2(x, out=None):
    _cuda(x)
    B, H, W, Cc = x.shape
    if out is None:
        out = torch.empty((B, H // 2, W // 2, Cc), device=x.device, dtype=x.dtype)
    check(lib.b200seg_maxpool2x2(ptr(x), ptr(out), _dt(x), B, H, W, Cc, _stream()), "maxpool2x2")
    return out


def softmax_ce(logits, target, want_grad: bool = True, grad_scale: Optional[float] = None):
    """Fused mean cross-entropy over NCHW f32 logits and int64 targets.
    Returns (loss scalar tensor, dlogits or None); dlogits already carries the 1/(B*H*W) factor."""
    _cuda(logits, target)
    if logits.dtype != torch.float32 or target.dtype != torch.int64:
        raise TypeError("softmax_ce expects f32 logits and int64 targets")
    B, Cc, H, W = logits.shape
    n = B * H * W
    loss_sum = torch.zeros(1, device=logits.device, dtype=torch.float32)
    dl = torch.empty_like(logits) if want_grad else None
    gs = (1.0 / n) if grad_scale is None else grad_scale
    check(lib.b200seg_softmax_ce(ptr(logits), ptr(target), ptr(loss_sum), ptr(dl), gs, B, Cc, H, W, _stream()),
          "softmax_ce")
    return loss_sum[0] / n, dl
