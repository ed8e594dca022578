"""T3: training path on the GPU.

Per-kernel: every backward / train-mode kernel against float64 autograd of the matching
torch.nn.functional op.  End-to-end: one fp32 training step of the drop-in MobileNetV2UNet / UNet
against the CPU oracle (loss, all 194 parameter gradients, BN running statistics,
num_batches_tracked, the untouched classifier) and two Adam steps through the reference's call
sequence (train.py:35-39) against the values frozen from the live reference.

Stated tolerances (fp32 path): loss 2e-5 absolute; logits 1e-4 max-relative; each gradient tensor within
max(1e-2, 10 x the reference's own fp32-vs-float64 deviation) of the float64 gradient, relative to the tensor
maximum (see _check_grads); gradients that are analytically zero are only required to be tiny.  bf16 path: loss within 2e-2, gradient cosine similarity > 0.8 on the decoder/outc.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

import b200seg  # noqa: E402
from b200seg import ops  # noqa: E402
from oracle import unet_oracle as O  # noqa: E402
from util import expand_aliases, fixture_sd, gold  # noqa: E402

DEV = "cuda"
TOL = {torch.float32: 3e-5, torch.bfloat16: 8e-3}


def _err(got, ref):
    got, ref = got.double(), ref.double()
    return float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def _rand(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,C,act,res", [(2, 16, 24, 32, 2, False), (3, 9, 7, 96, 1, False), (2, 8, 8, 24, 0, True),
                                             (1, 33, 65, 144, 2, False),
                                             # bf16: the cluster kernels (bn_cluster.cu) -- slice resident in shared memory,
                                             # cluster sizes 8 / 4, ragged pixel split, residual; streamed forward (128 ch)
                                             (8, 16, 32, 384, 2, False), (32, 8, 16, 960, 1, True), (5, 17, 31, 640, 0, False),
                                             (32, 32, 64, 128, 2, False)])
def test_batchnorm_train_forward_backward(dt, B, H, W, C, act, res):
    if dt == torch.bfloat16 and C >= 128 and C != 144:
        from b200seg._cabi import lib, BF16
        assert lib.b200seg_bn_cluster_supported(BF16, B * H * W, C) == 1        # these shapes must take the cluster path
        assert lib.b200seg_bn_cluster_bwd_supported(BF16, B * H * W, C) == (0 if C == 128 else 1)
    z = (_rand(B, C, H, W, seed=1) * 2 + 0.5).to(dt)
    gamma, beta = _rand(C, seed=2).abs() + 0.5, _rand(C, seed=3, scale=0.2)
    rm, rv = _rand(C, seed=4, scale=0.1), _rand(C, seed=5).abs() + 0.5
    r = _rand(B, C, H, W, seed=6).to(dt) if res else None
    da = _rand(B, C, H, W, seed=7).to(dt)
    # reference in float64
    zz = z.double().requires_grad_(True)
    g64, b64 = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rm64, rv64 = rm.double().clone(), rv.double().clone()
    y = F.batch_norm(zz, rm64, rv64, g64, b64, True, 0.1, 1e-5)
    y = {0: y, 1: F.relu(y), 2: torch.clamp(y, 0, 6)}[act]
    if res:
        y = y + r.double()
    y.backward(da.double())
    rm_k, rv_k = rm.clone(), rv.clone()
    a, sv = ops.bn_train_forward(_nhwc(z), gamma, beta, rm_k, rv_k, 1e-5, 0.1, act, _nhwc(r) if res else None)
    assert _err(_nchw(a), y.detach()) < TOL[dt]
    assert torch.allclose(rm_k.double(), rm64, rtol=1e-5, atol=1e-6) and torch.allclose(rv_k.double(), rv64, rtol=1e-5, atol=1e-6)
    dz, dgamma, dbeta = ops.bn_train_backward(_nhwc(da), _nhwc(z), sv, act)
    # a clamp boundary hit exactly by a bf16 value may flip one mask element: compare in the rms sense for bf16
    tol = 2e-4 if dt == torch.float32 else 2e-2
    assert _err(_nchw(dz), zz.grad) < tol
    assert _err(dgamma, g64.grad) < tol and _err(dbeta, b64.grad) < tol


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,C,act,res", [(2, 24, 40, 96, 2, False), (3, 17, 23, 24, 1, True), (2, 64, 128, 32, 2, False),
                                             (1, 9, 11, 2056, 0, False)])
def test_batchnorm_constants_fused_into_the_apply_kernels(dt, B, H, W, C, act, res, monkeypatch):
    """The streamed BatchNorm path of the training step derives the per-channel constants inside the apply kernels
    (b200seg_bn_finalize_apply, b200seg_bn_bwd_apply_slots).  Same arithmetic as the bn_finalize / f64->f32 launches they
    replace, so the results agree to rounding -- not bit for bit: the f64 slot sums themselves depend on the order in which
    the reduction blocks' atomics land (C = 2056: two channel blocks)."""
    monkeypatch.setattr(ops, "BN_CLUSTER", False)
    z = _nhwc((_rand(B, C, H, W, seed=31) * 2 + 0.5).to(dt))
    gamma, beta = _rand(C, seed=32).abs() + 0.5, _rand(C, seed=33, scale=0.2)
    r = _nhwc(_rand(B, C, H, W, seed=34).to(dt)) if res else None
    da = _nhwc(_rand(B, C, H, W, seed=35).to(dt))
    out = {}
    for fused in (False, True):
        monkeypatch.setattr(ops, "BN_FUSED_CONST", fused)
        rm, rv = _rand(C, seed=36, scale=0.1), _rand(C, seed=37).abs() + 0.5
        a, sv = ops.bn_train_forward(z, gamma, beta, rm, rv, 1e-5, 0.1, act, r)
        red = torch.zeros(ops.NSLOT, 2, C, device=DEV, dtype=torch.float64)
        dz, _, _ = ops.bn_train_backward(da, z, sv, act, red=red)
        out[fused] = (a, sv, rm, rv, dz, red)
    (a0, sv0, rm0, rv0, dz0, red0), (a1, sv1, rm1, rv1, dz1, red1) = out[False], out[True]
    for x0, x1 in ((sv0, sv1), (rm0, rm1), (rv0, rv1), (red0, red1)):
        assert torch.allclose(x0.double(), x1.double(), rtol=1e-5, atol=1e-6)
    for x0, x1 in ((a0, a1), (dz0, dz1)):
        d = (x0.float() - x1.float()).abs()
        scale = x0.float().abs().max()
        assert float(d.max()) <= (1e-5 if dt == torch.float32 else 8e-3) * float(scale)        # at most one ulp of the storage type
        if dt == torch.bfloat16:      # fp32: one channel whose constant sits on a rounding edge changes all its elements by an ulp
            assert float((d > 0).float().mean()) < 1e-3


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,Cin,Cout,taps", [(2, 16, 32, 64, 64, 1), (2, 9, 13, 24, 144, 1), (1, 16, 16, 16, 16, 1),
                                                 (2, 16, 32, 32, 32, 9), (1, 11, 7, 152, 64, 9), (3, 8, 16, 80, 40, 9)])
def test_conv_wgrad_and_dgrad(dt, B, H, W, Cin, Cout, taps):
    k = 3 if taps == 9 else 1
    x = _rand(B, Cin, H, W, seed=8).to(dt)
    w = _rand(Cout, Cin, k, k, seed=9, scale=(2.0 / (Cin * taps)) ** 0.5)
    dz = _rand(B, Cout, H, W, seed=10).to(dt)
    xx, ww = x.double().requires_grad_(True), w.double().requires_grad_(True)
    F.conv2d(xx, ww, None, 1, k // 2).backward(dz.double())
    dwk = ops.conv_wgrad(_nhwc(x), _nhwc(dz), taps)
    got_w = dwk.reshape(Cout, k, k, Cin).permute(0, 3, 1, 2)
    assert _err(got_w, ww.grad) < (1e-4 if dt == torch.float32 else 1e-3)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, -1)
    wt = wk.reshape(Cout, k, k, Cin).flip(1, 2).permute(3, 1, 2, 0).reshape(Cin, -1).contiguous()
    acc = _rand(B, Cin, H, W, seed=11).to(dt)
    got_x = ops.conv_simt(_nhwc(dz), wt, None, taps, 0, _nhwc(acc))
    assert _err(_nchw(got_x), xx.grad + acc.double()) < TOL[dt]
    if dt == torch.bfloat16 and Cin % 8 == 0 and Cout % 8 == 0:
        got_tc = ops.conv_tc(_nhwc(dz), wt.bfloat16(), None, taps, 0, _nhwc(acc))
        ref_tc = F.conv_transpose2d(dz.double(), w.bfloat16().double(), None, 1, k // 2) + acc.double()
        assert _err(_nchw(got_tc), ref_tc) < TOL[dt]


@pytest.mark.parametrize("B,H,W,Cin,Cout,taps", [
    (2, 16, 32, 64, 64, 1), (2, 16, 32, 64, 128, 1), (4, 32, 64, 16, 96, 1), (2, 16, 16, 144, 24, 1), (2, 8, 16, 320, 1280, 1),
    (1, 7, 9, 40, 72, 1), (2, 16, 32, 64, 64, 9), (2, 16, 32, 1344, 256, 9), (1, 32, 64, 288, 128, 9), (2, 64, 128, 80, 32, 9),
    (1, 23, 40, 64, 64, 9), (2, 5, 7, 24, 40, 9), (4, 128, 256, 32, 16, 1),
    # rows >= 128 px: tap-row mode (one CTA per dh, three accumulators, x tile with a 1-pixel halo)
    (1, 32, 256, 32, 32, 9), (1, 20, 200, 48, 40, 9), (1, 16, 160, 200, 64, 9), (1, 9, 130, 16, 136, 9)])
def test_conv_wgrad_tensor_core(B, H, W, Cin, Cout, taps):
    """tcgen05 weight gradient (MN-major operands, pixels as the reduction axis) vs float64 autograd."""
    k = 3 if taps == 9 else 1
    x = _rand(B, Cin, H, W, seed=8).bfloat16()
    dz = _rand(B, Cout, H, W, seed=10).bfloat16()
    ww = torch.zeros(Cout, Cin, k, k, device=DEV, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), ww, None, 1, k // 2).backward(dz.double())
    got = ops.conv_wgrad_tc(_nhwc(x), _nhwc(dz), taps).reshape(Cout, k, k, Cin).permute(0, 3, 1, 2)
    torch.cuda.synchronize()
    e = _err(got, ww.grad)
    if e >= 1e-3:
        d = (got.double() - ww.grad).abs()
        bad = (d > 1e-3 * ww.grad.abs().max()).nonzero()
        raise AssertionError(f"wgrad_tc err {e:.3e}; {bad.shape[0]} bad of {d.numel()}; first (co,ci,kh,kw) {bad[:6].tolist()} last {bad[-3:].tolist()}; "
                             f"got {got[tuple(bad[0])].item():.4f} ref {ww.grad[tuple(bad[0])].item():.4f}")


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,W,C,stride", [(2, 16, 32, 32, 1), (2, 16, 32, 96, 2), (1, 9, 7, 24, 2), (1, 7, 9, 144, 1),
                                            (2, 8, 16, 960, 1)])
def test_depthwise_backward(dt, B, H, W, C, stride):
    x = _rand(B, C, H, W, seed=12).to(dt)
    w = _rand(C, 1, 3, 3, seed=13, scale=0.4).bfloat16().float()     # bf16-representable taps (the bf16 kernels round them)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    dz = _rand(B, C, Ho, Wo, seed=14).to(dt)
    xx, ww = x.double().requires_grad_(True), w.double().requires_grad_(True)
    F.conv2d(xx, ww, None, stride, 1, 1, C).backward(dz.double())
    w9 = w.reshape(C, 9).t().contiguous()
    acc = _rand(B, C, H, W, seed=15).to(dt)
    dx = ops.dw_dgrad(_nhwc(dz), w9, (B, H, W, C), stride, _nhwc(acc))
    assert _err(_nchw(dx), xx.grad + acc.double()) < TOL[dt]
    dw9 = ops.dw_wgrad(_nhwc(x), _nhwc(dz), stride)
    assert _err(dw9.t().reshape(C, 1, 3, 3), ww.grad) < (1e-4 if dt == torch.float32 else 1e-3)


@pytest.mark.parametrize("xdt,dt", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16)])
@pytest.mark.parametrize("B,H,W,Cout,stride", [(2, 32, 48, 32, 2), (1, 17, 19, 16, 1), (1, 16, 16, 64, 1),
                                               (3, 40, 300, 32, 1), (2, 37, 517, 32, 2), (2, 21, 260, 64, 1), (5, 96, 512, 32, 2)])
def test_smallcin_wgrad(xdt, dt, B, H, W, Cout, stride):
    x = _rand(B, 3, H, W, seed=16).to(xdt)
    w = _rand(Cout, 3, 3, 3, seed=17, scale=0.3)
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    dz = _rand(B, Cout, Ho, Wo, seed=18).to(dt)
    ww = w.double().requires_grad_(True)
    F.conv2d(x.double(), ww, None, stride, 1).backward(dz.double())
    got = ops.smallcin_wgrad(x, _nhwc(dz), stride).permute(3, 2, 0, 1)
    assert _err(got, ww.grad) < 1e-4


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,h,w,Cs,Cu", [(2, 8, 16, 64, 128), (1, 3, 5, 8, 16), (1, 1, 1, 8, 8), (2, 16, 32, 16, 64)])
def test_upcat_backward(dt, B, h, w, Cs, Cu):
    x = _rand(B, Cu, h, w, seed=19)
    dcat = _rand(B, Cs + Cu, 2 * h, 2 * w, seed=20).to(dt)
    xx = x.double().requires_grad_(True)
    F.interpolate(xx, scale_factor=2, mode="bilinear", align_corners=False).backward(dcat[:, Cs:].double())
    acc = _rand(B, Cs, 2 * h, 2 * w, seed=21).to(dt)
    dskip, dx = ops.upcat_bwd(_nhwc(dcat), Cs, _nhwc(acc))
    assert _err(_nchw(dx), xx.grad) < TOL[dt]
    assert _err(_nchw(dskip), dcat[:, :Cs].double() + acc.double()) < TOL[dt]


@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,h,w,C", [(2, 16, 24, 10), (1, 5, 7, 3), (1, 1, 2, 16), (1, 64, 128, 10)])
def test_final_upsample_backward(dt, B, h, w, C):
    lg = _rand(B, C, h, w, seed=22).double().requires_grad_(True)
    dout = _rand(B, C, 2 * h, 2 * w, seed=23)
    F.interpolate(lg, scale_factor=2, mode="bilinear", align_corners=True).backward(dout.double())
    got = ops.final_bwd(dout, dt)
    assert got.shape == (B, h, w, 16)
    assert _err(_nchw(got)[:, :C], lg.grad) < TOL[dt]
    assert float(got[..., C:].abs().max()) == 0.0 if C < 16 else True


def test_maxpool_backward_and_colsum():
    for dt in (torch.float32, torch.bfloat16):
        x = _rand(2, 16, 12, 20, seed=24).to(dt)
        dy = _rand(2, 16, 6, 10, seed=25).to(dt)
        xx = x.double().requires_grad_(True)
        F.max_pool2d(xx, 2).backward(dy.double())
        got = ops.maxpool_bwd(_nhwc(x), _nhwc(dy))
        assert _err(_nchw(got), xx.grad) < 1e-6
        assert _err(ops.colsum(_nhwc(x)), x.double().sum(dim=(0, 2, 3))) < (1e-5 if dt == torch.float32 else 1e-5)


# ------------------------------------------------------------------------------------------------
# end to end
# ------------------------------------------------------------------------------------------------
def _oracle_step(sd, x, t, forward=O.mobilenetv2_unet_forward, dt=torch.float32):
    params = {k: (v.to(dt) if v.is_floating_point() else v).clone().requires_grad_(v.is_floating_point() and "running" not in k)
              for k, v in sd.items()}
    upd = O.BNState()
    out = forward(params, x.to(dt), training=True, upd=upd)
    loss = O.cross_entropy(out, t)
    loss.backward()
    return out.detach(), loss.detach(), params, upd


def _check_grads(model, params32, params64):
    """Gradients against the float64 oracle.  This network's fp32 gradients are ill-conditioned (ReLU6 masks and
    BatchNorm-backward cancellations on as few as 8 samples per channel at 1/32 scale): the REFERENCE's own fp32
    CPU gradients deviate from float64 by 0.1-5 % of the tensor maximum, per tensor, because a 1e-6 forward
    difference flips a few ReLU6 masks.  Stated tolerance: each gradient tensor within max(1e-2, 10 x the fp32
    oracle's own deviation) of the float64 gradient relative to the tensor maximum, AND cosine similarity
    >= 0.9999; analytically-zero gradients (shifts removed by a following train-mode BN) must be tiny."""
    report = {}
    for name, p in model.named_parameters():
        ref64 = params64[name].grad
        if name.startswith("backbone.classifier"):
            assert p.grad is None and ref64 is None          # SURVEY finding 5
            continue
        assert p.grad is not None, name
        got = p.grad.detach().cpu().double()
        scale = float(ref64.abs().max())
        if scale < 1e-5:
            assert float(got.abs().max()) < 1e-4, name
            continue
        floor = float((params32[name].grad.double() - ref64).abs().max()) / scale
        err = float((got - ref64).abs().max()) / scale
        cos = float((got.flatten() @ ref64.flatten()) / (got.norm() * ref64.norm()).clamp_min(1e-300))
        report[name] = (err, floor, cos)
    bad = {k: v for k, v in report.items() if v[0] > max(1e-2, 10 * v[1]) or v[2] < 0.9999}
    assert not bad, sorted(bad.items(), key=lambda kv: -kv[1][0])[:8]
    return report


def test_mbv2unet_fp32_train_step_matches_oracle():
    sd = fixture_sd()
    x, t = O.synth_input(2, 64, 64, seed=1), O.synth_target(2, 64, 64, seed=1)
    ref_out, ref_loss, params, upd = _oracle_step(sd, x, t)
    _, _, params64, _ = _oracle_step(sd, x, t, dt=torch.float64)
    m = b200seg.MobileNetV2UNet(output_channels=10)
    m.load_state_dict(expand_aliases(sd), strict=True)
    m = m.to(DEV).train()
    out = m(x.to(DEV))
    loss = b200seg.CrossEntropyLoss()(out, t.to(DEV))
    loss.backward()
    assert out.dtype == torch.float32 and out.shape == (2, 10, 64, 64)
    assert _err(out.detach().cpu(), ref_out) < 1e-4
    assert abs(float(loss) - float(ref_loss)) < 2e-5
    frozen = gold("mbv2unet_train.npz")
    assert abs(float(loss) - float(frozen["loss"])) < 2e-5          # the live reference's own loss
    rep = _check_grads(m, params, params64)
    import json, os
    worst = sorted(rep.items(), key=lambda kv: -kv[1][0])[:5]
    os.makedirs(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out"), exist_ok=True)
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_notes.jsonl"), "a") as f:
        f.write(json.dumps(dict(test="fp32_train_grads", n=len(rep), median_err=sorted(v[0] for v in rep.values())[len(rep) // 2],
                                median_floor=sorted(v[1] for v in rep.values())[len(rep) // 2], worst=worst)) + "\n")
    new_sd = m.state_dict()
    for k, v in upd.updates.items():                                  # BN running stats + num_batches_tracked
        got = new_sd[k].cpu()
        if v.dtype == torch.long:
            assert int(got) == int(v), k
        else:
            assert torch.allclose(got, v, rtol=2e-4, atol=2e-6), k
    # eval after train uses the freshly updated running statistics (packed-weight cache invalidation)
    m.eval()
    with torch.no_grad():
        ye = m(x.to(DEV)).cpu()
        sd2 = dict(sd); sd2.update(upd.updates)
        assert _err(ye, O.mobilenetv2_unet_forward(sd2, x)) < 1e-4


def test_two_adam_steps_through_reference_call_sequence():
    """train.py:35-39 verbatim: zero_grad / forward / criterion / backward / step, torch.optim.Adam(lr=1.5e-4)."""
    frozen = gold("mbv2unet_train.npz")
    sd = fixture_sd()
    m = b200seg.MobileNetV2UNet(output_channels=10)
    m.load_state_dict(expand_aliases(sd), strict=True)
    m = m.to(DEV)
    criterion = torch.nn.CrossEntropyLoss()                       # the stock loss also works (main.py:99)
    optimizer = torch.optim.Adam(m.parameters(), lr=1.5e-4)       # main.py:100
    m.train()
    losses = []
    for seed in (1, 2):
        inputs, targets = O.synth_input(2, 64, 64, seed=seed).to(DEV), O.synth_target(2, 64, 64, seed=seed).to(DEV)
        optimizer.zero_grad()
        outputs = m(inputs)
        loss = criterion(outputs, targets)
        loss.backward()
        optimizer.step()
        losses.append(loss.item())
    assert abs(losses[0] - float(frozen["step_losses"][0])) < 2e-5
    assert abs(losses[1] - float(frozen["step_losses"][1])) < 2e-4
    new_sd = m.state_dict()
    for k in frozen.files:
        if not k.startswith("p2:") or k == "p2:up4.conv.conv.3.bias":   # zero-gradient bias: Adam amplifies noise
            continue
        ref = torch.from_numpy(frozen[k]).double()
        d = (new_sd[k[3:]].cpu().double() - ref).abs()
        # Adam normalises each element's gradient: where |g| is at the fp32 noise floor the step direction itself is
        # noise, so a few elements may differ by up to one lr per step; the bulk must agree far below one step.
        # (at step 1 every element moves by exactly +-lr whatever |g| is; at step 2 the move depends on g2/g1, so an
        # element whose gradient is at the noise floor can be off by a large fraction of lr.)
        assert float(d.max()) <= 4 * 1.5e-4 + 1e-7, (k, float(d.max()))      # 2 steps x (at most +-lr each)
        assert float(d.mean()) < 0.15 * 1.5e-4, (k, float(d.mean()))
        assert float((d > 0.75e-4).double().mean()) < 0.05, (k, float((d > 0.75e-4).double().mean()))
    assert int(new_sd["outc.conv.1.num_batches_tracked"]) == 2
    assert m.backbone.classifier[1].weight.grad is None


def test_plain_unet_fp32_train_step_matches_oracle():
    sd = O.synth_state_dict(O.unet_param_shapes(10, 16), seed=3)
    x, t = O.synth_input(1, 32, 48, seed=3), O.synth_target(1, 32, 48, seed=3)
    ref_out, ref_loss, params, upd = _oracle_step(sd, x, t, forward=O.unet_forward)
    _, _, params64, _ = _oracle_step(sd, x, t, forward=O.unet_forward, dt=torch.float64)
    m = b200seg.UNet(output_channels=10, base_filters=16)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).train()
    out = m(x.to(DEV))
    loss = F.cross_entropy(out, t.to(DEV))
    loss.backward()
    assert _err(out.detach().cpu(), ref_out) < 1e-4
    assert abs(float(loss) - float(ref_loss)) < 2e-5
    assert abs(float(loss) - float(gold("unet_eval.npz")["loss"])) < 2e-5
    _check_grads(m, params, params64)


def test_bf16_training_path_tracks_fp32():
    """bf16 activations + tcgen05 convs (fp32 master weights): loss and gradients stay close to fp32.
    (A 4x3x128x128 batch: BatchNorm over a handful of samples, as at 1/32 scale of a 64x64 crop, is too
    ill-conditioned in bf16 to say anything.)"""
    sd = fixture_sd()
    x, t = O.synth_input(4, 128, 128, seed=1), O.synth_target(4, 128, 128, seed=1)
    _, ref_loss, params, _ = _oracle_step(sd, x, t)
    m = b200seg.MobileNetV2UNet(output_channels=10)
    m.load_state_dict(expand_aliases(sd), strict=True)
    m = m.to(DEV).train()
    m._get_engine().precision = "bf16"
    out = m(x.to(DEV))
    loss = b200seg.CrossEntropyLoss()(out, t.to(DEV))
    loss.backward()
    assert abs(float(loss.detach()) - float(ref_loss)) < 2e-2
    cosines = {}
    for name, p in m.named_parameters():
        if p.grad is None or params[name].grad is None or p.grad.numel() < 64:
            continue
        a, b = p.grad.detach().cpu().flatten().double(), params[name].grad.flatten().double()
        if float(b.abs().max()) < 1e-5:
            continue
        cosines[name] = float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))
    # the floor: the oracle graph run by eager PyTorch in bf16 on this GPU (what `.bfloat16()` training of the
    # reference would give)
    p16 = {k: (v.to(DEV).bfloat16() if v.is_floating_point() else v.to(DEV)).requires_grad_(v.is_floating_point() and "running" not in k)
           for k, v in sd.items()}
    o16 = O.mobilenetv2_unet_forward(p16, x.to(DEV).bfloat16(), training=True, upd=O.BNState())
    F.cross_entropy(o16.float(), t.to(DEV)).backward()
    floor = {}
    for name in cosines:
        a, b = p16[name].grad.detach().float().cpu().flatten().double(), params[name].grad.flatten().double()
        floor[name] = float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))
    import json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    vals, fvals = sorted(cosines.values()), sorted(floor.values())
    with open(os.path.join(root, "gpurun_out", "parity_notes.jsonl"), "a") as f:
        f.write(json.dumps(dict(test="bf16_train_grad_cosine", n=len(vals), min=vals[0], p10=vals[len(vals) // 10],
                                median=vals[len(vals) // 2], eager_bf16_floor=dict(min=fvals[0], p10=fvals[len(fvals) // 10], median=fvals[len(fvals) // 2]),
                                worst=[(k, v, floor[k]) for k, v in sorted(cosines.items(), key=lambda kv: kv[1])[:6]])) + "\n")
    assert vals[len(vals) // 2] > fvals[len(fvals) // 2] - 0.1, (vals[len(vals) // 2], fvals[len(fvals) // 2])
    for name in ("outc.conv.3.weight", "outc.conv.0.weight", "up4.conv.conv.3.weight"):
        assert cosines[name] > min(0.9, floor[name] - 0.05), (name, cosines[name], floor[name])


def test_train_step_cuda_graph_replay_matches_eager():
    """The CUDA-graph replay computes the same step as the kernel-by-kernel launches.  Compared after ONE step from
    identical state (graph captured on the first call): the loss and every BatchNorm running statistic agree to
    float rounding (atomic summation order), parameters up to the summation order of the float atomics in the
    weight-gradient kernels.  A longer run then checks that replays keep advancing the state."""
    sd = fixture_sd()

    def make(use_graphs, graph_after):
        m = b200seg.MobileNetV2UNet(output_channels=10)
        m.load_state_dict(expand_aliases(sd), strict=True)
        m = m.to(DEV).train()
        eng = m._get_engine()
        eng.use_graphs, eng.graph_after = use_graphs, graph_after
        return m, torch.optim.Adam(m.parameters(), lr=1.5e-4), b200seg.CrossEntropyLoss()

    def one(m, opt, crit, seed):
        x, t = O.synth_input(2, 64, 64, seed=seed).to(DEV), O.synth_target(2, 64, 64, seed=seed).to(DEV)
        opt.zero_grad()
        loss = crit(m(x), t)
        loss.backward()
        opt.step()
        return loss.item()

    me, oe, ce = make(False, 2)
    mg, og, cg = make(True, 0)
    le, lg = one(me, oe, ce, 30), one(mg, og, cg, 30)
    assert any(getattr(e.get("graph"), "bwd", None) is not None for e in mg._get_engine()._graphs.values()), "not captured"
    # BatchNorm statistics are accumulated with float64 atomics whose order differs launch to launch: ~1e-7 relative
    assert abs(le - lg) <= 2e-6 * abs(le), (le, lg)
    sde, sdg = me.state_dict(), mg.state_dict()
    for k in sde:
        if "num_batches" in k:
            assert torch.equal(sde[k], sdg[k]), k
        elif "running" in k:
            assert torch.allclose(sde[k], sdg[k], rtol=1e-5, atol=1e-7), k
    for k in ("outc.conv.3.weight", "up1.conv.conv.0.weight", "backbone.features.0.0.weight", "backbone.features.18.1.weight"):
        d = (sde[k] - sdg[k]).abs()
        assert float(d.max()) <= 2 * 1.5e-4 + 1e-7 and float(d.mean()) < 1e-6, (k, float(d.max()), float(d.mean()))
    losses = [one(mg, og, cg, 31 + i) for i in range(5)]
    assert all(l == l and l < 10 for l in losses)
    assert int(mg.state_dict()["outc.conv.1.num_batches_tracked"]) == 6


# ---------------------------------------------------------------------------------------------------------------
# fused multi-tensor Adam against torch.optim.Adam
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("wd", [0.0, 1e-2])
def test_fused_adam_matches_torch_adam(wd):
    """Same update rule as torch/optim/adam.py on tensors of awkward sizes (unaligned views, < 4 elements, > 1 chunk),
    5 steps with fresh gradients; state_dict layouts are interchangeable."""
    g = torch.Generator(device="cpu").manual_seed(3)
    shapes = [(1,), (3,), (7, 5), (32, 3, 3, 3), (96, 16, 1, 1), (1280, 320, 1, 1), (10,), (4099,)]
    base = [torch.randn(*s, generator=g).to(DEV) for s in shapes]
    big = torch.randn(1000 + 3, generator=g).to(DEV)
    base.append(big[3:])                                          # 12-byte offset: not 16-byte aligned -> scalar path
    pa = [b.clone().requires_grad_(True) for b in base[:-1]] + [base[-1].clone().requires_grad_(True)]
    pb = [b.clone().requires_grad_(True) for b in base[:-1]] + [big.clone()[3:].requires_grad_(True)]
    oa = torch.optim.Adam(pa, lr=1.5e-4, weight_decay=wd)
    ob = b200seg.Adam(pb, lr=1.5e-4, weight_decay=wd)
    for step in range(5):
        for x, y in zip(pa, pb):
            gr = torch.randn(x.shape, generator=g).to(DEV) * (10.0 ** (step - 2))
            x.grad = gr.clone()
            y.grad = gr.clone()
        oa.step()
        ob.step()
    for x, y in zip(pa, pb):
        d = (x.detach() - y.detach()).abs().max().item()
        assert d <= 1e-6, d                                        # a few fp32 ulps of O(1) parameters; the update itself is 7.5e-4
    sa, sb = oa.state_dict(), ob.state_dict()
    assert sa["state"].keys() == sb["state"].keys()
    for k in sa["state"]:
        assert set(sa["state"][k].keys()) == set(sb["state"][k].keys()) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"]) == 5.0
        torch.testing.assert_close(sa["state"][k]["exp_avg"], sb["state"][k]["exp_avg"], rtol=1e-5, atol=1e-12)
        torch.testing.assert_close(sa["state"][k]["exp_avg_sq"], sb["state"][k]["exp_avg_sq"], rtol=1e-5, atol=1e-20)
    ob2 = b200seg.Adam(pb, lr=1.5e-4, weight_decay=wd)
    ob2.load_state_dict(sa)                                       # torch's optimizer state resumes in the fused one


def test_fused_adam_trains_the_model_like_torch_adam():
    """Two steps of the reference loop body with b200seg.Adam vs torch.optim.Adam from identical state (fp32 path)."""
    sd = fixture_sd()

    def run(opt_cls):
        m = b200seg.MobileNetV2UNet(output_channels=10)
        m.load_state_dict(expand_aliases(sd), strict=True)
        m = m.to(DEV).train()
        m._get_engine().use_graphs = False
        opt, crit = opt_cls(m.parameters(), lr=1.5e-4), b200seg.CrossEntropyLoss()
        for seed in (40, 41):
            x, t = O.synth_input(2, 64, 64, seed=seed).to(DEV), O.synth_target(2, 64, 64, seed=seed).to(DEV)
            opt.zero_grad()
            loss = crit(m(x), t)
            loss.backward()
            opt.step()
        return {k: v.detach().clone() for k, v in m.named_parameters()}

    a, b = run(torch.optim.Adam), run(b200seg.Adam)
    for k in a:
        d = (a[k] - b[k]).abs()
        # identical gradients up to atomics order; Adam turns rounding noise of ~zero gradients into +-lr steps, and
        # the 9 conv biases in front of a train-mode BatchNorm have an analytically zero gradient (pure noise): for
        # those only the step size bounds the difference, run to run, with ANY optimizer implementation
        zero_grad_bias = k.endswith(".bias") and (".conv.conv.0." in k or ".conv.conv.3." in k or k.startswith("outc.conv.0."))
        assert float(d.max()) <= 4 * 1.5e-4, (k, float(d.max()))
        if not zero_grad_bias:
            # the two runs differ by the atomics order of the reductions; at steps 1-2 Adam's update is ~lr*sign(g), so
            # the handful of near-zero gradients that flip sign move by up to 2*lr (exact arithmetic agreement of the two
            # optimizers on identical gradients is test_fused_adam_matches_torch_adam)
            assert float(d.mean()) < 0.5 * 1.5e-4, (k, float(d.max()), float(d.mean()))
