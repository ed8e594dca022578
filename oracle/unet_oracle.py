"""CPU oracle for the segmentation hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch fp32 *functional* restatement of the reference's
forward path (``/root/reference/src/unet.py`` plus the un-vendored torchvision
``mobilenet_v2`` encoder it instantiates).  It exists so that the CUDA path can
be checked on a box where ``/root/reference`` is absent.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it; the product package (``b200seg``) never does and
has no CPU fallback.

Parity pinning: the reference ships no tests or golden vectors, so
this restatement is pinned against the *live* reference imported in the authoring
container by ``oracle/make_golden.py``; the resulting vectors are committed under
``tests/golden/`` and re-checked by ``tests/test_oracle_golden.py`` on every run.

The state_dict consumed here uses the reference's own key names (SURVEY.md
Appendix A): ``backbone.features.N...``, ``upK.conv.conv.{0,1,3,4}...``,
``outc.conv.{0,1,3}...``.  Third-party dependency restated: torchvision 0.26.0
``models/mobilenetv2.py:19-64,101-161`` and ``ops/misc.py:69-126`` (unpinned in
the reference's requirements.txt:2).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5        # nn.BatchNorm2d default, used by every BN in unet.py / torchvision
BN_MOMENTUM = 0.1

# torchvision mobilenetv2.py:105-114 -- (expand t, out channels c, repeats n, first stride s)
MBV2_SETTING = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2),
                (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]


def mbv2_blocks() -> List[Tuple[int, int, int, int, int]]:
    """(feature index, inp, oup, stride, expand) for features[1..17]
    (torchvision mobilenetv2.py:129-136)."""
    out, inp, idx = [], 32, 1
    for t, c, n, s in MBV2_SETTING:
        for i in range(n):
            out.append((idx, inp, c, s if i == 0 else 1, t))
            inp, idx = c, idx + 1
    return out


class BNState:
    """Side channel that collects running-stat updates in train mode."""

    def __init__(self):
        self.updates: Dict[str, Tensor] = {}
        self.batch_stats: Dict[str, Tuple[Tensor, Tensor]] = {}


def _bn(sd, prefix: str, x: Tensor, training: bool, upd: Optional[BNState]) -> Tensor:
    """nn.BatchNorm2d forward.  train: batch mean / biased var
    over (B,H,W); running stats take momentum 0.1 with the *unbiased* variance."""
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    if not training:
        rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
        return F.batch_norm(x, rm, rv, w, b, False, 0.0, BN_EPS)
    # nn.BatchNorm2d.forward in training mode == F.batch_norm(..., training=True, momentum=0.1) on the
    # module's own running buffers; here the buffers are clones so the caller's dict is not mutated.
    rm = sd[prefix + ".running_mean"].detach().clone()
    rv = sd[prefix + ".running_var"].detach().clone()
    y = F.batch_norm(x, rm, rv, w, b, True, BN_MOMENTUM, BN_EPS)
    if upd is not None:
        with torch.no_grad():
            upd.updates[prefix + ".running_mean"] = rm
            upd.updates[prefix + ".running_var"] = rv
            upd.updates[prefix + ".num_batches_tracked"] = sd[prefix + ".num_batches_tracked"] + 1
            upd.batch_stats[prefix] = (x.detach().mean(dim=(0, 2, 3)), x.detach().var(dim=(0, 2, 3), unbiased=False))
    return y


def _relu6(x):  # torchvision ops/misc.py:114 (ReLU6 inplace) == clamp(x, 0, 6)
    return torch.clamp(x, 0.0, 6.0)


def _conv_bn_act(sd, p: str, x, stride, groups, training, upd, pad):
    """torchvision Conv2dNormActivation (ops/misc.py:86-116): conv(bias=False)->BN->ReLU6."""
    x = F.conv2d(x, sd[p + ".0.weight"], None, stride, pad, 1, groups)
    return _relu6(_bn(sd, p + ".1", x, training, upd))


def _inverted_residual(sd, p: str, x, inp, oup, stride, expand, training, upd):
    """torchvision mobilenetv2.py:19-64."""
    hidden = int(round(inp * expand))
    y, i = x, 0
    if expand != 1:
        y = _conv_bn_act(sd, f"{p}.conv.{i}", y, 1, 1, training, upd, 0); i += 1
    y = _conv_bn_act(sd, f"{p}.conv.{i}", y, stride, hidden, training, upd, 1); i += 1
    y = F.conv2d(y, sd[f"{p}.conv.{i}.weight"]); i += 1          # pw-linear, no bias
    y = _bn(sd, f"{p}.conv.{i}", y, training, upd)
    if stride == 1 and inp == oup:                               # mobilenetv2.py:32,61-62
        y = x + y
    return y


def _double_conv(sd, p: str, x, training, upd):
    """unet.py:53-68: (conv3x3 pad1 +bias -> BN -> ReLU) x 2."""
    x = F.conv2d(x, sd[p + ".conv.0.weight"], sd[p + ".conv.0.bias"], 1, 1)
    x = F.relu(_bn(sd, p + ".conv.1", x, training, upd))
    x = F.conv2d(x, sd[p + ".conv.3.weight"], sd[p + ".conv.3.bias"], 1, 1)
    return F.relu(_bn(sd, p + ".conv.4", x, training, upd))


def _up(sd, p: str, x1, x2, training, upd):
    """unet.py:94-105: bilinear x2 (align_corners=False, unet.py:97), cat([skip, up]) (unet.py:103)."""
    x1 = F.interpolate(x1, scale_factor=2, mode="bilinear", align_corners=False)
    return _double_conv(sd, p + ".conv", torch.cat([x2, x1], dim=1), training, upd)


def _outconv(sd, p: str, x, training, upd):
    """unet.py:108-121: 1x1+b -> BN -> ReLU -> 1x1+b."""
    x = F.conv2d(x, sd[p + ".conv.0.weight"], sd[p + ".conv.0.bias"])
    x = F.relu(_bn(sd, p + ".conv.1", x, training, upd))
    return F.conv2d(x, sd[p + ".conv.3.weight"], sd[p + ".conv.3.bias"])


def mobilenetv2_unet_forward(sd: Dict[str, Tensor], x: Tensor, training: bool = False,
                             upd: Optional[BNState] = None, taps: Optional[dict] = None) -> Tensor:
    """MobileNetV2UNet.forward (unet.py:32-51).  ``taps`` (optional dict) receives the
    intermediate feature maps x1..x5 and decoder outputs for per-stage parity tests."""
    f = "backbone.features"
    y = _conv_bn_act(sd, f + ".0", x, 2, 1, training, upd, 1)    # stem, mobilenetv2.py:125-127
    skips = {}
    for idx, inp, oup, stride, expand in mbv2_blocks():
        y = _inverted_residual(sd, f"{f}.{idx}", y, inp, oup, stride, expand, training, upd)
        if idx in (1, 3, 6, 10):                                 # unet.py:15-18 slice ends
            skips[idx] = y
    y = _conv_bn_act(sd, f + ".18", y, 1, 1, training, upd, 0)   # 320->1280, mobilenetv2.py:139-143
    x1, x2, x3, x4, x5 = skips[1], skips[3], skips[6], skips[10], y
    u1 = _up(sd, "up1", x5, x4, training, upd)
    u2 = _up(sd, "up2", u1, x3, training, upd)
    u3 = _up(sd, "up3", u2, x2, training, upd)
    u4 = _up(sd, "up4", u3, x1, training, upd)
    logits_half = _outconv(sd, "outc", u4, training, upd)
    out = F.interpolate(logits_half, scale_factor=2, mode="bilinear", align_corners=True)  # unet.py:30,49
    if taps is not None:
        taps.update(x1=x1, x2=x2, x3=x3, x4=x4, x5=x5, u1=u1, u2=u2, u3=u3, u4=u4,
                    logits_half=logits_half)
    return out


def unet_forward(sd: Dict[str, Tensor], x: Tensor, training: bool = False,
                 upd: Optional[BNState] = None) -> Tensor:
    """UNet.forward (unet.py:137-147); inconv :71-78, down :81-91 (MaxPool2d(2) + double_conv)."""
    x1 = _double_conv(sd, "inc.conv", x, training, upd)
    x2 = _double_conv(sd, "down1.mpconv.1", F.max_pool2d(x1, 2), training, upd)
    x3 = _double_conv(sd, "down2.mpconv.1", F.max_pool2d(x2, 2), training, upd)
    x4 = _double_conv(sd, "down3.mpconv.1", F.max_pool2d(x3, 2), training, upd)
    y = _up(sd, "up1", x4, x3, training, upd)
    y = _up(sd, "up2", y, x2, training, upd)
    y = _up(sd, "up3", y, x1, training, upd)
    return _outconv(sd, "sem_out", y, training, upd)


def cross_entropy(logits: Tensor, target: Tensor) -> Tensor:
    """nn.CrossEntropyLoss() defaults (main.py:99, train.py:37): mean over B*H*W pixels of
    -log_softmax(logits)[target]; log-sum-exp with max subtraction."""
    m = logits.max(dim=1, keepdim=True).values
    lse = (logits - m).exp().sum(dim=1, keepdim=True).log() + m
    picked = torch.gather(logits, 1, target[:, None]).to(logits.dtype)
    return (lse - picked).mean()


def cross_entropy_grad(logits: Tensor, target: Tensor) -> Tensor:
    """d loss / d logits = (softmax - onehot) / (B*H*W)."""
    p = torch.softmax(logits, dim=1)
    onehot = torch.zeros_like(p).scatter_(1, target[:, None], 1.0)
    return (p - onehot) / (logits.numel() // logits.shape[1])


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float = 1.5e-4,
              b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam single-tensor update with the defaults main.py:100 relies on
    (no weight decay, no amsgrad).  In place on p, m, v.  ``step`` is 1-based."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


# --------------------------------------------------------------------------------------
# Deterministic fixtures that do NOT depend on constructor RNG order: every tensor is
# drawn from its own seeded generator, so the authoring container (which applies them to
# the live reference) and the GPU box (which applies them to the CUDA path) agree.
# --------------------------------------------------------------------------------------
def mbv2unet_param_shapes(out_ch: int = 10) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(key, shape, kind) for every *distinct* tensor of MobileNetV2UNet in
    ``named_parameters``/buffer order under the ``backbone.*`` / ``up*`` / ``outc`` spelling.
    kind in {conv, dw, bn_w, bn_b, bn_rm, bn_rv, bn_n, bias, fc_w, fc_b}."""
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def bn(p, c):
        out.extend([(p + ".weight", (c,), "bn_w"), (p + ".bias", (c,), "bn_b"),
                    (p + ".running_mean", (c,), "bn_rm"), (p + ".running_var", (c,), "bn_rv"),
                    (p + ".num_batches_tracked", (), "bn_n")])

    f = "backbone.features"
    out.append((f + ".0.0.weight", (32, 3, 3, 3), "conv")); bn(f + ".0.1", 32)
    for idx, inp, oup, stride, expand in mbv2_blocks():
        hidden, i = inp * expand, 0
        if expand != 1:
            out.append((f"{f}.{idx}.conv.0.0.weight", (hidden, inp, 1, 1), "conv"))
            bn(f"{f}.{idx}.conv.0.1", hidden); i = 1
        out.append((f"{f}.{idx}.conv.{i}.0.weight", (hidden, 1, 3, 3), "dw"))
        bn(f"{f}.{idx}.conv.{i}.1", hidden)
        out.append((f"{f}.{idx}.conv.{i + 1}.weight", (oup, hidden, 1, 1), "conv"))
        bn(f"{f}.{idx}.conv.{i + 2}", oup)
    out.append((f + ".18.0.weight", (1280, 320, 1, 1), "conv")); bn(f + ".18.1", 1280)
    out.append(("backbone.classifier.1.weight", (1000, 1280), "fc_w"))
    out.append(("backbone.classifier.1.bias", (1000,), "fc_b"))
    for name, cin, cout in (("up1", 1344, 256), ("up2", 288, 128), ("up3", 152, 64), ("up4", 80, 32)):
        p = f"{name}.conv.conv"
        out.append((p + ".0.weight", (cout, cin, 3, 3), "conv")); out.append((p + ".0.bias", (cout,), "bias"))
        bn(p + ".1", cout)
        out.append((p + ".3.weight", (cout, cout, 3, 3), "conv")); out.append((p + ".3.bias", (cout,), "bias"))
        bn(p + ".4", cout)
    out.append(("outc.conv.0.weight", (16, 32, 1, 1), "conv")); out.append(("outc.conv.0.bias", (16,), "bias"))
    bn("outc.conv.1", 16)
    out.append(("outc.conv.3.weight", (out_ch, 16, 1, 1), "conv")); out.append(("outc.conv.3.bias", (out_ch,), "bias"))
    return out


def unet_param_shapes(out_ch: int = 10, base: int = 64) -> List[Tuple[str, Tuple[int, ...], str]]:
    out: List[Tuple[str, Tuple[int, ...], str]] = []

    def bn(p, c):
        out.extend([(p + ".weight", (c,), "bn_w"), (p + ".bias", (c,), "bn_b"),
                    (p + ".running_mean", (c,), "bn_rm"), (p + ".running_var", (c,), "bn_rv"),
                    (p + ".num_batches_tracked", (), "bn_n")])

    def dc(p, cin, cout):
        out.append((p + ".0.weight", (cout, cin, 3, 3), "conv")); out.append((p + ".0.bias", (cout,), "bias"))
        bn(p + ".1", cout)
        out.append((p + ".3.weight", (cout, cout, 3, 3), "conv")); out.append((p + ".3.bias", (cout,), "bias"))
        bn(p + ".4", cout)

    b = base
    dc("inc.conv.conv", 3, b)
    dc("down1.mpconv.1.conv", b, 2 * b); dc("down2.mpconv.1.conv", 2 * b, 4 * b); dc("down3.mpconv.1.conv", 4 * b, 4 * b)
    dc("up1.conv.conv", 8 * b, 2 * b); dc("up2.conv.conv", 4 * b, b); dc("up3.conv.conv", 2 * b, b)
    out.append(("sem_out.conv.0.weight", (b // 2, b, 1, 1), "conv")); out.append(("sem_out.conv.0.bias", (b // 2,), "bias"))
    bn("sem_out.conv.1", b // 2)
    out.append(("sem_out.conv.3.weight", (out_ch, b // 2, 1, 1), "conv")); out.append(("sem_out.conv.3.bias", (out_ch,), "bias"))
    return out


def synth_state_dict(shapes, seed: int = 0) -> Dict[str, Tensor]:
    """Seed-per-tensor synthetic weights in a *realistic trained-like* regime:
    conv ~ N(0, 2/fan_in) (He), BN gamma ~ U(0.5,1.5), beta ~ N(0,0.1), running_mean ~ N(0,0.1),
    running_var ~ U(0.5,1.5).  Keeps activations O(1) through all 62 convs so that parity
    tolerances are meaningful (SURVEY finding 10: raw random init in eval() is degenerate)."""
    sd: Dict[str, Tensor] = {}
    for i, (key, shape, kind) in enumerate(shapes):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        if kind in ("conv", "dw"):
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif kind == "fc_w":
            t = torch.randn(shape, generator=g) * 0.01
        elif kind in ("bn_w", "bn_rv"):
            t = torch.rand(shape, generator=g) + 0.5
        elif kind in ("bn_b", "bn_rm", "bias", "fc_b"):
            t = torch.randn(shape, generator=g) * 0.1
        elif kind == "bn_n":
            t = torch.zeros((), dtype=torch.long)
        else:
            raise ValueError(kind)
        sd[key] = t
    return sd


def synth_input(b: int, h: int, w: int, seed: int = 0) -> Tensor:
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.randn(b, 3, h, w, generator=g)


def synth_target(b: int, h: int, w: int, ncls: int = 10, seed: int = 0) -> Tensor:
    """Blocky road-scene-like labels: class 0 sky/top, class 1 road/bottom, rectangles of 2..9."""
    g = torch.Generator().manual_seed(2000 + seed)
    t = torch.zeros(b, h, w, dtype=torch.long)
    t[:, h // 2:, :] = 1
    for i in range(b):
        for _ in range(6):
            c = int(torch.randint(2, ncls, (1,), generator=g))
            y0 = int(torch.randint(0, h - 4, (1,), generator=g)); x0 = int(torch.randint(0, w - 4, (1,), generator=g))
            hh = int(torch.randint(4, max(5, h // 3), (1,), generator=g)); ww = int(torch.randint(4, max(5, w // 3), (1,), generator=g))
            t[i, y0:y0 + hh, x0:x0 + ww] = c
    return t


def calibrate_bn(sd: Dict[str, Tensor], x: Tensor, forward=None) -> Dict[str, Tensor]:
    """Fixture F1: replace every BN's running stats by the batch statistics of one
    train-mode pass over ``x`` (what ``momentum=None`` after one batch would give, biased var).
    Returns a new dict; used only to put synthetic weights into a realistic eval regime."""
    upd = BNState()
    with torch.no_grad():
        (forward or mobilenetv2_unet_forward)(sd, x, training=True, upd=upd)
    out = dict(sd)
    for prefix, (mean, var) in upd.batch_stats.items():
        out[prefix + ".running_mean"] = mean
        out[prefix + ".running_var"] = var
    return out


def bn_stat_keys(sd) -> List[str]:
    return [k for k in sd if k.endswith("running_mean") or k.endswith("running_var")]


# --------------------------------------------------------------------------------------
# Fixture F2: a TRAINED network.  The road-scene generator below is what the live
# reference is trained on by oracle/make_golden_f2.py; the weights it ends with are committed
# (int8 per output channel, see f2_state_dict) so that every box evaluates the same network.
# --------------------------------------------------------------------------------------
ROAD_PALETTE = [(0.45, 0.65, 0.95), (0.30, 0.30, 0.32), (0.85, 0.15, 0.15), (0.15, 0.75, 0.20), (0.95, 0.85, 0.10),
                (0.55, 0.25, 0.70), (0.95, 0.55, 0.10), (0.10, 0.80, 0.80), (0.90, 0.90, 0.90), (0.05, 0.05, 0.05)]
ROAD_MEAN, ROAD_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)       # inference.py:36-37


def road_scene_batch(b: int, h: int, w: int, seed: int = 0, noise: float = 0.15) -> Tuple[Tensor, Tensor]:
    """Synthetic 10-class road scene: class 0 on the top half, class 1 (road) on the bottom half, six
    random rectangles of classes 2..9 per image; the image is the class palette colour + N(0, noise^2), normalised
    with the ImageNet statistics the reference uses.  numpy PCG64 streams: identical on every machine."""
    import numpy as np
    rng = np.random.default_rng(770000 + seed)
    t = np.zeros((b, h, w), dtype=np.int64)
    t[:, h // 2:, :] = 1
    for i in range(b):
        for _ in range(6):
            c = int(rng.integers(2, 10))
            y0, x0 = int(rng.integers(0, h - 4)), int(rng.integers(0, w - 4))
            hh, ww = int(rng.integers(4, max(5, h // 3))), int(rng.integers(4, max(5, w // 3)))
            t[i, y0:y0 + hh, x0:x0 + ww] = c
    pal = np.asarray(ROAD_PALETTE, dtype=np.float32)
    img = pal[t] + rng.standard_normal((b, h, w, 3), dtype=np.float32) * np.float32(noise)
    img = (img - np.asarray(ROAD_MEAN, np.float32)) / np.asarray(ROAD_STD, np.float32)
    return torch.from_numpy(np.ascontiguousarray(img.transpose(0, 3, 1, 2))), torch.from_numpy(t)


def f2_quantize(sd: Dict[str, Tensor]) -> Dict[str, "object"]:
    """Trained state_dict -> committed form: conv/linear weights as int8 with one f32 scale per output channel,
    everything else (BN vectors, biases, running stats) as f32.  f2_state_dict() inverts it exactly."""
    import numpy as np
    out = {}
    for k, v in sd.items():
        if v.dim() == 4 and v.is_floating_point():
            w = v.detach().float()
            s = w.abs().amax(dim=(1, 2, 3)).clamp_min(1e-12) / 127.0
            q = torch.round(w / s[:, None, None, None]).clamp_(-127, 127).to(torch.int8)
            out["q:" + k] = q.numpy()
            out["s:" + k] = s.numpy()
        elif v.is_floating_point():
            out["f:" + k] = v.detach().float().numpy()
        else:
            out["i:" + k] = v.detach().numpy()
    return out


def f2_state_dict(blob) -> Dict[str, Tensor]:
    """The F2 network every box evaluates: W = scale[o] * int8 (one exact f32 product per element)."""
    import numpy as np
    sd: Dict[str, Tensor] = {}
    for key in blob.files if hasattr(blob, "files") else blob.keys():
        kind, name = key[:2], key[2:]
        arr = torch.from_numpy(np.ascontiguousarray(blob[key]))
        if kind == "q:":
            sd[name] = arr.float() * torch.from_numpy(np.ascontiguousarray(blob["s:" + name]))[:, None, None, None]
        elif kind in ("f:", "i:"):
            sd[name] = arr
    return sd
