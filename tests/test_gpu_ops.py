"""T1: every CUDA kernel against the matching torch.nn.functional op (computed in
float64 on the same GPU, i.e. no TF32 anywhere) on the reference's real layer shapes plus ragged
ones.  All calls go through the C ABI (b200seg.ops -> ctypes -> libb200seg.so).

Tolerances: f32 storage: 2e-5 of the output range.  bf16 storage: inputs/weights are rounded to
bf16 first and the reference is computed from the rounded values, so the only error left is fp32
accumulation order + one bf16 output rounding (2^-9 relative) -> 6e-3 of the output range.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from b200seg import ops  # noqa: E402

DEV = "cuda"
TOL = {torch.float32: 2e-5, torch.bfloat16: 6e-3}


def _err(got, ref):
    got, ref = got.double(), ref.double()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def _nhwc(t):  # NCHW -> NHWC contiguous
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def _act(y, act):
    return {0: y, 1: F.relu(y), 2: torch.clamp(y, 0, 6)}[act]


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("xdt,ydt", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                     (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("B,H,W,Cout,stride,act", [(2, 64, 96, 32, 2, 2), (1, 33, 47, 32, 2, 2), (1, 32, 48, 64, 1, 1),
                                                   (1, 17, 19, 16, 1, 0)])
def test_conv3x3_smallcin(xdt, ydt, B, H, W, Cout, stride, act):
    x = _rand(B, 3, H, W, seed=1).to(xdt)
    w = _rand(Cout, 3, 3, 3, seed=2, scale=0.3)
    b = _rand(Cout, seed=3, scale=0.1)
    ref = _act(F.conv2d(x.double(), w.double(), b.double(), stride, 1), act)
    got = ops.conv3x3_smallcin(x, w.permute(2, 3, 1, 0).contiguous(), b, stride, act, ydt)
    assert got.shape == (B, ref.shape[2], ref.shape[3], Cout)
    assert _err(_nchw(got), ref) < TOL[ydt]


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,C,stride", [(2, 128, 256, 32, 1), (2, 128, 256, 96, 2), (1, 64, 128, 144, 1),
                                            (2, 16, 32, 576, 2), (3, 8, 16, 960, 1), (1, 7, 9, 24, 1), (1, 9, 7, 8, 2),
                                            (1, 1, 1, 16, 1), (1, 23, 40, 192, 2)])
def test_dwconv3x3(dt, B, H, W, C, stride):
    x = _rand(B, C, H, W, seed=4).to(dt)
    w = _rand(C, 1, 3, 3, seed=5, scale=0.4)
    b = _rand(C, seed=6, scale=0.1)
    ref = torch.clamp(F.conv2d(x.double(), w.double(), b.double(), stride, 1, 1, C), 0, 6)
    got = ops.dwconv3x3(_nhwc(x), w.reshape(C, 9).t().contiguous(), b, stride, 2)
    assert _err(_nchw(got), ref) < TOL[dt]


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("B,H,W,C,stride", [(2, 128, 256, 32, 1), (2, 128, 256, 96, 2), (1, 64, 128, 144, 1),
                                            (2, 16, 32, 576, 2), (3, 8, 16, 960, 1), (1, 7, 9, 24, 1), (1, 9, 7, 8, 2),
                                            (1, 1, 1, 16, 1), (1, 23, 40, 192, 2)])
def test_dwconv3x3_bf16_taps_mixed_precision_fma(variant, B, H, W, C, stride):
    """bf16 activations x bf16 taps on the mixed-precision FMA: equals the float64 convolution of the same bf16 operands up
    to fp32 accumulation + the output rounding, and equals the f32-tap kernel BIT FOR BIT (bf16 products are exact in f32)."""
    x = _rand(B, C, H, W, seed=4).bfloat16()
    w = _rand(C, 1, 3, 3, seed=5, scale=0.4).bfloat16()
    b = _rand(C, seed=6, scale=0.1)
    ref = torch.clamp(F.conv2d(x.double(), w.double(), b.double(), stride, 1, 1, C), 0, 6)
    w9c = w.reshape(C, 9).t().contiguous()
    got = ops.dwconv3x3_bf16w(_nhwc(x), w9c, b, stride, 2, variant=variant)
    assert _err(_nchw(got), ref) < TOL[torch.bfloat16]
    assert torch.equal(got, ops.dwconv3x3(_nhwc(x), w9c.float(), b, stride, 2))


@pytest.mark.parametrize("flags", [0, 2], ids=["auto", "tap_mode"])
@pytest.mark.parametrize("B,H,W,C,stride", [(2, 128, 256, 32, 1), (1, 128, 256, 96, 2), (1, 64, 128, 144, 1),
                                            (2, 16, 32, 576, 2), (3, 8, 16, 960, 1), (1, 7, 9, 24, 1), (1, 9, 7, 8, 2),
                                            (1, 1, 1, 16, 1), (1, 23, 40, 192, 2), (1, 33, 100, 72, 1), (1, 64, 128, 64, 2)])
def test_dwconv3x3_tensor_core(flags, B, H, W, C, stride):
    """Depthwise 3x3 run as a block-diagonal implicit GEMM on tcgen05 (weights are bf16 on this path)."""
    x = _rand(B, C, H, W, seed=4).bfloat16()
    w = _rand(C, 1, 3, 3, seed=5, scale=0.4).bfloat16()
    b = _rand(C, seed=6, scale=0.1)
    ref = torch.clamp(F.conv2d(x.double(), w.double(), b.double(), stride, 1, 1, C), 0, 6)
    wdiag = ops.pack_dw_diag(w.float().reshape(C, 9).t().contiguous())
    got = ops.dwconv3x3_tc(_nhwc(x), wdiag, b, stride, 2, flags=flags)
    torch.cuda.synchronize()
    e = _err(_nchw(got), ref)
    if e >= TOL[torch.bfloat16]:
        d = (_nchw(got).double() - ref).abs()
        bad = (d > TOL[torch.bfloat16] * ref.abs().max()).nonzero()
        raise AssertionError(f"dw_tc err {e:.3e}; {bad.shape[0]} bad of {d.numel()}; first (b,c,h,w) {bad[:6].tolist()} last {bad[-3:].tolist()}")


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,Cin,Cout,taps,act,res", [
    (2, 16, 32, 64, 64, 1, 2, False), (1, 8, 16, 320, 1280, 1, 2, False), (2, 16, 32, 384, 64, 1, 0, True),
    (1, 13, 9, 24, 144, 1, 2, False), (1, 16, 32, 64, 64, 9, 1, False), (1, 5, 7, 152, 64, 9, 1, False),
    (1, 8, 8, 16, 12, 1, 0, False)])
def test_conv_simt(dt, B, H, W, Cin, Cout, taps, act, res):
    k = 3 if taps == 9 else 1
    x = _rand(B, Cin, H, W, seed=7).to(dt)
    w = _rand(Cout, Cin, k, k, seed=8, scale=(2.0 / (Cin * taps)) ** 0.5)
    b = _rand(Cout, seed=9, scale=0.1)
    r = _rand(B, Cout, H, W, seed=10).to(dt) if res else None
    ref = _act(F.conv2d(x.double(), w.double(), b.double(), 1, k // 2), act)
    if res:
        ref = ref + r.double()
    got = ops.conv_simt(_nhwc(x), w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous(), b, taps, act,
                        _nhwc(r) if res else None)
    assert _err(_nchw(got), ref) < TOL[dt]


# (B,H,W,Cin,Cout,taps,act,res): the real MobileNetV2UNet layers at a small batch + ragged cases
TC_CASES = [
    (1, 8, 16, 64, 64, 1, 0, False),        # one M tile, one K chunk, N=64: the minimal tcgen05 case
    (2, 16, 32, 64, 128, 1, 2, False),
    (2, 128, 256, 32, 16, 1, 0, False),     # features.1 project
    (1, 128, 256, 16, 96, 1, 2, False),     # features.2 expand (K=16: one UMMA k-step)
    (1, 64, 128, 96, 24, 1, 0, False),      # N=24 -> UMMA N 32, store clipped
    (2, 64, 128, 24, 144, 1, 2, False),     # K=24: ragged k chunk
    (2, 64, 128, 144, 24, 1, 0, True),      # residual
    (2, 16, 32, 384, 64, 1, 0, True),
    (2, 16, 32, 64, 384, 1, 2, False),      # 2 N tiles of 192
    (2, 8, 16, 160, 960, 1, 2, False),      # 5 N tiles of 192
    (2, 8, 16, 960, 320, 1, 0, False),      # 15 K chunks, N tiles 192+128(clipped)
    (2, 8, 16, 320, 1280, 1, 2, False),     # features.18
    (1, 128, 256, 32, 16, 1, 1, False),     # outc.0
    (1, 128, 256, 16, 16, 1, 0, False),     # outc.3 (padded to 16)
    (1, 7, 9, 40, 72, 1, 2, True),          # ragged M (63 pixels), ragged K and N
    (2, 16, 32, 64, 64, 9, 1, False),       # 3x3, tile = 4 rows x 32
    (2, 16, 32, 1344, 256, 9, 1, False),    # up1.conv.0
    (1, 32, 64, 288, 128, 9, 1, False),     # up2.conv.0 (ragged K chunk)
    (1, 64, 128, 152, 64, 9, 1, False),     # up3.conv.0
    (1, 128, 256, 80, 32, 9, 1, False),     # up4.conv.0 (tile = half a row)
    (1, 128, 256, 32, 32, 9, 1, False),
    (1, 23, 40, 64, 64, 9, 1, False),       # 736x1280 / 32: tile 3 x 40 = 120 rows
    (2, 5, 7, 24, 40, 9, 0, True),          # tiny ragged everything
    (1, 46, 80, 64, 32, 9, 1, False),
]


@pytest.mark.parametrize("flags", [0, 1], ids=["direct_store", "tma_store"])
@pytest.mark.parametrize("B,H,W,Cin,Cout,taps,act,res", TC_CASES)
def test_conv_tc(flags, B, H, W, Cin, Cout, taps, act, res):
    k = 3 if taps == 9 else 1
    x = _rand(B, Cin, H, W, seed=11).bfloat16()
    w = _rand(Cout, Cin, k, k, seed=12, scale=(2.0 / (Cin * taps)) ** 0.5).bfloat16()
    b = _rand(Cout, seed=13, scale=0.1)
    r = _rand(B, Cout, H, W, seed=14).bfloat16() if res else None
    ref = _act(F.conv2d(x.double(), w.double(), b.double(), 1, k // 2), act)
    if res:
        ref = ref + r.double()
    got = ops.conv_tc(_nhwc(x), w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous(), b, taps, act,
                      _nhwc(r) if res else None, flags=flags)
    torch.cuda.synchronize()
    e = _err(_nchw(got), ref)
    if e >= TOL[torch.bfloat16]:
        d = (_nchw(got).double() - ref).abs()
        bad = (d > TOL[torch.bfloat16] * ref.abs().max()).nonzero()
        raise AssertionError(f"conv_tc err {e:.3e}; {bad.shape[0]} bad of {d.numel()}; first bad (b,c,h,w) {bad[:5].tolist()} "
                             f"last {bad[-3:].tolist()}; got {_nchw(got)[tuple(bad[0])].item()} ref {ref[tuple(bad[0])].item()}")


def test_conv_tc_rejects_bad_arguments():
    x = torch.zeros(1, 4, 4, 12, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(16, 12, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        ops.conv_tc(x, w, None, 1, 0)
    with pytest.raises(TypeError):
        ops.conv_tc(x.float(), w, None, 1, 0)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,h,w,Cs,Cu", [(2, 8, 16, 64, 1280), (1, 16, 32, 32, 256), (1, 32, 64, 24, 128),
                                         (2, 64, 128, 16, 64), (1, 3, 5, 8, 16), (1, 1, 1, 8, 8)])
def test_upsample2x_concat(dt, B, h, w, Cs, Cu):
    x = _rand(B, Cu, h, w, seed=15).to(dt)
    skip = _rand(B, Cs, 2 * h, 2 * w, seed=16).to(dt)
    ref = torch.cat([skip.double(), F.interpolate(x.double(), scale_factor=2, mode="bilinear", align_corners=False)], 1)
    got = ops.upsample2x_concat(_nhwc(skip), _nhwc(x))
    assert _err(_nchw(got), ref) < TOL[dt]
    assert torch.equal(_nchw(got)[:, :Cs], skip)                 # the skip half is a bit-exact copy


@pytest.mark.parametrize("dt,odt", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16),
                                    (torch.bfloat16, torch.float32)])
@pytest.mark.parametrize("B,h,w,C", [(2, 128, 256, 10), (1, 16, 24, 10), (1, 5, 7, 3), (1, 1, 2, 16)])
def test_final_upsample_align_corners(dt, odt, B, h, w, C):
    lg = torch.zeros(B, h, w, 16, device=DEV, dtype=dt)
    lg[..., :C] = _nhwc(_rand(B, C, h, w, seed=17)).to(dt)
    lg[..., C:] = 1e9 if C < 16 else 0      # padding channels must never leak into the output
    ref = F.interpolate(_nchw(lg[..., :C].contiguous()).double(), scale_factor=2, mode="bilinear", align_corners=True)
    got = ops.upsample2x_ac_nchw(lg, C, odt)
    assert got.shape == (B, C, 2 * h, 2 * w) and got.dtype == odt
    assert _err(got, ref) < TOL[odt]
    if 2 * w % 4 == 0:
        mask = ops.upsample2x_ac_argmax(lg, C)
        ref32 = F.interpolate(_nchw(lg[..., :C].contiguous()).float(), scale_factor=2, mode="bilinear", align_corners=True)
        agree = (mask.long() == ref32.argmax(1)).float().mean().item()
        assert agree > 0.999, agree


def test_maxpool_and_nhwc_to_nchw():
    for dt in (torch.float32, torch.bfloat16):
        x = _rand(2, 64, 12, 20, seed=18).to(dt)
        got = ops.maxpool2x2(_nhwc(x))
        assert torch.equal(_nchw(got), F.max_pool2d(x, 2))
        y = ops.nhwc_to_nchw(_nhwc(x), 10, torch.float32)
        assert torch.equal(y, x[:, :10].float())


@pytest.mark.parametrize("B,C,H,W", [(2, 10, 64, 128), (1, 10, 7, 9), (3, 19, 8, 8), (1, 1, 4, 4), (2, 13, 16, 24), (3, 16, 4, 4)])
def test_softmax_ce_fused_forward_and_gradient(B, C, H, W):
    lg = _rand(B, C, H, W, seed=19, scale=3.0).requires_grad_(True)
    tg = torch.randint(0, C, (B, H, W), generator=torch.Generator().manual_seed(20)).to(DEV)
    ref = F.cross_entropy(lg.double(), tg)
    (gref,) = torch.autograd.grad(ref, lg)
    loss, dl = ops.softmax_ce(lg.detach(), tg)
    assert abs(loss.item() - ref.item()) < 2e-6 * max(1.0, abs(ref.item()))
    assert _err(dl, gref) < 1e-5
    # module form used at the train.py:37-38 call site
    import b200seg
    l2 = b200seg.CrossEntropyLoss()(lg, tg)
    l2.backward()
    assert _err(lg.grad, gref) < 1e-5


# ------------------------------------------------------------------------------------------------
# fused output tail: outconv(32, C) + final_upsample (+ argmax)  (unet.py:47-49, :108-121; inference.py:64)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,h,w,C", [(2, 16, 32, 10), (1, 24, 40, 1), (3, 48, 80, 16), (1, 368, 640, 10), (2, 16, 16, 7)])
def test_tail_fused_matches_the_three_ops_it_replaces(B, h, w, C):
    g = torch.Generator().manual_seed(h * 131 + C)
    x = torch.randn(B, h, w, 32, generator=g).bfloat16().cuda()
    w0 = (torch.randn(16, 32, generator=g) * 0.25).bfloat16().cuda()
    b0 = torch.zeros(64); b0[:16] = torch.randn(16, generator=g) * 0.3
    w3 = torch.zeros(16, 16); w3[:C] = torch.randn(C, 16, generator=g) * 0.35
    w3 = w3.bfloat16().cuda()
    b3 = torch.zeros(64); b3[:C] = torch.randn(C, generator=g) * 0.3
    b0, b3 = b0.cuda(), b3.cuda()
    hid = torch.relu(x.double() @ w0.double().t() + b0[:16].double()).bfloat16().double()      # bf16 where the unfused path stored it
    lg = (hid @ w3.double().t() + b3[:16].double())[..., :C].permute(0, 3, 1, 2)
    ref = F.interpolate(lg, scale_factor=2, mode="bilinear", align_corners=True)
    rng = float(ref.max() - ref.min()) + 1e-6
    y32 = ops.tail_fused(x, w0, b0, w3, b3, C, torch.float32)
    assert y32.shape == (B, C, 2 * h, 2 * w)
    # the hidden activation is rounded to bf16 from an fp32 accumulator here and from float64 in the reference: on large
    # tensors a few values sit on a rounding boundary and flip by one bf16 ulp (-> ~1e-3 of a logit); everything else agrees
    # to fp32 rounding
    d = (y32.double() - ref).abs().flatten() / rng
    assert float(d.max()) < 2e-3
    assert float(torch.quantile(d[:: max(1, d.numel() // 1000000)], 0.999)) < 2e-5
    y16 = ops.tail_fused(x, w0, b0, w3, b3, C, torch.bfloat16)
    assert float((y16.double() - ref).abs().max()) / rng < 5e-3
    mask = ops.tail_fused(x, w0, b0, w3, b3, C, want_mask=True)
    assert mask.dtype == torch.uint8 and mask.shape == (B, 2 * h, 2 * w)
    am = ref.argmax(1)
    top2 = ref.topk(min(2, C), dim=1).values
    tie = (top2[:, 0] - top2[:, -1]).abs() < 1e-4 * rng if C > 1 else torch.zeros_like(am, dtype=torch.bool)
    assert bool(((mask.long() == am) | tie).all())


# ------------------------------------------------------------------------------------------------
# row-stacked 3x3 conv for few output channels (conv_rs.cu): same contract as conv_tc(taps=9)
# ------------------------------------------------------------------------------------------------
RS_CASES = [  # B, H, W, Cin, Cout, act, res
    (2, 128, 256, 80, 32, 1, False),     # up4.conv.0: two column tiles per row, 2 K chunks (64 + 16)
    (2, 128, 256, 32, 32, 1, False),     # up4.conv.3
    (3, 40, 96, 32, 32, 0, True),        # narrow map (one 96-pixel tile), accumulate into res (the data-gradient use)
    (1, 3, 200, 8, 24, 2, False),        # minimal height, ragged second column tile, Cout 24 -> 32 columns, K = 8
    (5, 7, 130, 24, 16, 1, True),        # 16-column variant, many short strips (ranges span images)
    (1, 128, 256, 32, 8, 0, False),      # Cout = 8
    (2, 33, 257, 136, 32, 1, False),     # odd sizes, 3 K chunks
    (2, 128, 256, 32, 80, 0, True),      # data gradient of up4.conv.0 (32 -> 80): N = 240, accumulate into res
    (1, 9, 150, 16, 72, 1, False),       # 72 -> 80 columns, ragged tile
]


@pytest.mark.parametrize("flags", [0, 1 << 20, (2 << 16) | (1 << 20), (7 << 8)], ids=["auto", "1cta", "1cta_2stages", "grid28"])
@pytest.mark.parametrize("B,H,W,Cin,Cout,act,res", RS_CASES)
def test_conv_rs(flags, B, H, W, Cin, Cout, act, res):
    x = _rand(B, Cin, H, W, seed=21).bfloat16()
    w = _rand(Cout, Cin, 3, 3, seed=22, scale=(2.0 / (Cin * 9)) ** 0.5).bfloat16()
    b = _rand(Cout, seed=23, scale=0.1)
    r = _rand(B, Cout, H, W, seed=24).bfloat16() if res else None
    ref = _act(F.conv2d(x.double(), w.double(), b.double(), 1, 1), act)
    if res:
        ref = ref + r.double()
    wk = w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    got = ops.conv_rs(_nhwc(x), wk, b, act, _nhwc(r) if res else None, flags=flags)
    torch.cuda.synchronize()
    assert _err(_nchw(got), ref) < TOL[torch.bfloat16]
    # and against the tap-by-tap tensor-core kernel (flags=2: TAP addressing, never dispatched to conv_rs)
    old = ops.conv_tc(_nhwc(x), wk, b, 9, act, _nhwc(r) if res else None, flags=2)
    assert _err(got.float(), old.double()) < 8e-3          # one bf16 ulp of the largest output


def test_conv_tc_dispatches_small_cout_3x3_to_conv_rs():
    from b200seg._cabi import lib
    assert lib.b200seg_conv_rs_supported(128, 256, 80, 32) == 1 and lib.b200seg_conv_rs_supported(128, 256, 32, 32) == 1
    assert lib.b200seg_conv_rs_supported(64, 128, 152, 64) == 0          # 64 output channels: 3 accumulators exceed TMEM
    assert lib.b200seg_conv_rs_supported(32, 64, 32, 32) == 0            # narrow map: conv_tc's 2-D tiles
    assert lib.b200seg_conv_rs_supported(128, 256, 1344, 32) == 0        # weights would not stay resident


# ------------------------------------------------------------------------------------------------
# fused head: features.0 (stem 3x3 s2 + ReLU6) -> features.1 (depthwise 3x3 + ReLU6 -> linear 1x1) (stem_mb1.cu)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("xdt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
@pytest.mark.parametrize("B,H,W", [(2, 64, 128), (1, 16, 64), (3, 34, 136), (1, 2, 8), (2, 256, 512)])
def test_stem_mb1_matches_the_three_ops_it_replaces(B, H, W, xdt):
    x = _rand(B, 3, H, W, seed=31).to(xdt)
    w0 = _rand(32, 3, 3, 3, seed=32, scale=(2.0 / 27) ** 0.5)
    b0 = _rand(32, seed=33, scale=0.3)
    wd = _rand(32, 1, 3, 3, seed=34, scale=0.4).bfloat16().float()
    bd = _rand(32, seed=35, scale=0.3)
    wp = _rand(16, 32, seed=36, scale=(1.0 / 32) ** 0.5).bfloat16()
    bp = _rand(16, seed=37, scale=0.2)
    # float64 reference with the kernel's rounding points: bf16 input and stem weights, both intermediates rounded to bf16
    e = torch.clamp(F.conv2d(x.bfloat16().double(), w0.bfloat16().double(), b0.double(), 2, 1), 0, 6).bfloat16().double()
    d = torch.clamp(F.conv2d(e, wd.double(), bd.double(), 1, 1, 1, 32), 0, 6).bfloat16().double()
    ref = F.conv2d(d, wp.double()[:, :, None, None], bp.double())
    w0k = w0.permute(2, 3, 1, 0).contiguous()                      # [kh][kw][c][32]
    wdk = wd.reshape(32, 9).t().contiguous().bfloat16()            # [9][32]
    got = ops.stem_mb1(x, w0k, b0, wdk, bd, wp, bp)
    torch.cuda.synchronize()
    assert got.shape == (B, H // 2, W // 2, 16)
    assert _err(_nchw(got), ref) < TOL[torch.bfloat16]
    # and against the three kernels it replaces (same rounding points: equal up to fp32 summation order)
    s0 = ops.conv3x3_smallcin(x, w0k, b0, 2, 2, torch.bfloat16)
    s1 = ops.dwconv3x3_bf16w(s0, wdk, bd, 1, 2)
    s2 = ops.conv_tc(s1, wp, bp, 1, 0)
    assert _err(got.float(), s2.double()) < 8e-3
