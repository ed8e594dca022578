"""Fused inverted-residual kernel (b200seg_mbconv) against the three torch.nn.functional ops it replaces
(tv:models/mobilenetv2.py:38-62), computed in float64 from the same bf16-rounded operands with the two intermediate
activations rounded to bf16 exactly where the unfused path stores them.  Remaining error: fp32 accumulation order
and one bf16 rounding of the output (plus the rare intermediate that rounds the other way) -> 1e-2 of the range.
Shapes: every (Cin, Ce, Cout, stride) of the MobileNetV2 encoder, ragged maps, and the 720p map sizes.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from b200seg import ops  # noqa: E402

DEV = "cuda"


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def _case(B, H, W, Cin, Ce, Cout, stride, residual, seed=0):
    x = _rand(B, H, W, Cin, seed=seed + 1).bfloat16()
    we = _rand(Ce, Cin, seed=seed + 2, scale=(2.0 / Cin) ** 0.5).bfloat16()
    be = _rand(Ce, seed=seed + 3, scale=0.3)
    wd = _rand(9, Ce, seed=seed + 4, scale=0.4).bfloat16().float()      # bf16-representable taps (the kernel reads bf16)
    bd = _rand(Ce, seed=seed + 5, scale=0.3)
    wp = _rand(Cout, Ce, seed=seed + 6, scale=(1.0 / Ce) ** 0.5).bfloat16()
    bp = _rand(Cout, seed=seed + 7, scale=0.2)
    return x, we, be, wd, bd, wp, bp


def _reference(x, we, be, wd, bd, wp, bp, stride, residual):
    xd = x.double().permute(0, 3, 1, 2)
    Ce = we.shape[0]
    e = torch.clamp(F.conv2d(xd, we.double()[:, :, None, None], be.double()), 0, 6).bfloat16().double()
    d = torch.clamp(F.conv2d(e, wd.double().t().reshape(Ce, 1, 3, 3), bd.double(), stride, 1, 1, Ce), 0, 6)
    d = d.bfloat16().double()
    y = F.conv2d(d, wp.double()[:, :, None, None], bp.double())
    if residual:
        y = y + xd
    return y.permute(0, 2, 3, 1)


def _run(x, we, be, wd, bd, wp, bp, stride, residual, flags=0):
    return ops.mbconv(x, we, ops.pad_channels(be, 64), ops.pad_channels(wd.bfloat16(), 64), ops.pad_channels(bd, 64), wp,
                      ops.pad_channels(bp, 16), stride, residual, flags=flags)


CASES = [  # B, H, W, Cin, Ce, Cout, stride, residual      (encoder blocks features.2 .. features.17)
    (2, 32, 64, 16, 96, 24, 2, False),
    (2, 16, 32, 24, 144, 24, 1, True),
    (2, 32, 32, 24, 144, 32, 2, False),
    (2, 16, 32, 32, 192, 32, 1, True),
    (2, 16, 32, 32, 192, 64, 2, False),
    (2, 16, 32, 64, 384, 64, 1, True),
    (2, 16, 32, 64, 384, 96, 1, False),
    (2, 16, 32, 96, 576, 96, 1, True),
    (2, 16, 32, 96, 576, 160, 2, False),
    (3, 8, 16, 160, 960, 160, 1, True),
    (3, 8, 16, 160, 960, 320, 1, False),
    (1, 23, 40, 32, 192, 32, 1, True),       # ragged tiles (736x1280 input: 23x40 at 1/32)
    (1, 23, 40, 32, 192, 64, 2, False),
    (1, 5, 7, 16, 96, 24, 1, False),         # smaller than one tile
    (1, 9, 21, 24, 144, 24, 2, False),
    (5, 24, 48, 24, 144, 24, 1, True),       # many tiles per CTA
]


@pytest.mark.parametrize("B,H,W,Cin,Ce,Cout,stride,residual", CASES)
def test_mbconv_matches_unfused_reference(B, H, W, Cin, Ce, Cout, stride, residual):
    ops_in = _case(B, H, W, Cin, Ce, Cout, stride, residual)
    ref = _reference(*ops_in, stride, residual)
    got = _run(*ops_in, stride, residual)
    assert got.shape == ref.shape
    err = float((got.double() - ref).abs().max() / ref.abs().max())
    assert err < 1e-2, err


@pytest.mark.parametrize("flags", [1 | (1 << 2) | (1 << 4), 2 | (2 << 2) | (2 << 4) | (1 << 6), 1 | (2 << 2) | (1 << 4),
                                   (1 << 8), 1 << 16, 2 << 16, (1 << 6) | (2 << 16), (1 << 6) | (1 << 16)])
def test_mbconv_buffering_variants_agree(flags):
    """Single/double buffered rings, one or two CTAs per SM, 8- or 4-row tiles and a 4-CTA grid (many tiles per CTA)
    give bit-identical results."""
    args = _case(4, 32, 64, 32, 192, 32, 1, True, seed=10)
    base = _run(*args, 1, True)
    assert torch.equal(_run(*args, 1, True, flags=flags), base)


def test_mbconv_deterministic_and_repeatable():
    args = _case(8, 32, 64, 64, 384, 64, 1, True, seed=20)
    a = _run(*args, 1, True)
    for _ in range(3):
        assert torch.equal(_run(*args, 1, True), a)


def test_mbconv_rejects_bad_arguments():
    x, we, be, wd, bd, wp, bp = _case(1, 8, 16, 16, 96, 24, 1, False)
    with pytest.raises(RuntimeError):
        _run(x, we, be, wd, bd, wp, bp, 3, False)                 # stride
    with pytest.raises(RuntimeError):
        _run(x, we, be, wd, bd, wp, bp, 1, True)                  # residual with Cin != Cout
    with pytest.raises((ValueError, RuntimeError)):
        ops.mbconv(x, we, be, wd.bfloat16(), bd, wp, bp, 1, False)   # unpadded parameter vectors
    with pytest.raises(TypeError):
        _run(x, we, be, wd, bd, wp.float(), bp, 1, False)         # f32 weights
