"""Generate tests/golden/* by running the LIVE reference (/root/reference) -- authoring
container only; TEST INFRASTRUCTURE, never imported by the product.

The reference is Python and cannot travel to the GPU box, so its outputs on seeded
synthetic weights/inputs are frozen here as small fixtures:

  tests/golden/mbv2unet_keys.json      691 state_dict keys (order, shape, dtype) + alias groups
                                       + named_parameters order           (unet.py:8-30)
  tests/golden/unet_keys.json          same for UNet(10)                  (unet.py:124-135)
  tests/golden/mbv2unet_bnstats.npz    BN running stats of fixture F1 (calibrated once here, frozen)
  tests/golden/mbv2unet_eval.npz       eval logits, taps (x1..x5,u1..u4) checksums   (unet.py:32-51)
  tests/golden/mbv2unet_train.npz      train-mode loss, logits, selected grads, BN running stats,
                                       params after 2 Adam steps via the reference's own
                                       train_model()                      (train.py:6-79, main.py:99-100)
  tests/golden/unet_eval.npz           UNet(10, base 16) eval/train logits + loss

Run:  python oracle/make_golden.py          (needs /root/reference; ~1 min on 8 vCPU)
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def load_reference():
    """SURVEY Appendix D shim: unet.py:12 asks for pretrained weights (a download);
    force weights=None before the constructor runs."""
    import torchvision.models as tvm
    _orig = tvm.mobilenet_v2
    tvm.mobilenet_v2 = lambda *a, weights=None, **k: _orig(*a, weights=None, **k)
    sys.path.insert(0, "/root/reference")
    from src import unet as ref_unet
    from src import train as ref_train
    return ref_unet, ref_train


def expand_aliases(sd):
    """backbone.features.N.* also appears as downK.N.* (unet.py:15-19)."""
    out = dict(sd)
    for k, v in sd.items():
        if k.startswith("backbone.features."):
            n = int(k.split(".")[2])
            d = 1 if n < 2 else 2 if n < 4 else 3 if n < 7 else 4 if n < 11 else 5
            out[f"down{d}." + k[len("backbone.features."):]] = v
    return out


def keys_json(model, path):
    sd = model.state_dict()
    ptr = {}
    for k, v in sd.items():
        ptr.setdefault(v.data_ptr() if v.numel() else ("n", k), []).append(k)
    info = {
        "keys": [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()],
        "alias_groups": [g for g in ptr.values() if len(g) > 1],
        "named_parameters": [n for n, _ in model.named_parameters()],
        "named_buffers": [n for n, _ in model.named_buffers()],
        "n_params": sum(p.numel() for p in model.parameters()),
    }
    with open(path, "w") as f:
        json.dump(info, f)
    print(path, len(info["keys"]), "keys", len(info["alias_groups"]), "alias groups")


def main():
    torch.set_num_threads(8)
    os.makedirs(GOLD, exist_ok=True)
    ref_unet, ref_train = load_reference()

    # ---------------- structure ----------------
    torch.manual_seed(0)
    ref = ref_unet.MobileNetV2UNet(output_channels=10)
    keys_json(ref, os.path.join(GOLD, "mbv2unet_keys.json"))
    keys_json(ref_unet.UNet(output_channels=10), os.path.join(GOLD, "unet_keys.json"))

    # ---------------- MobileNetV2UNet eval ----------------
    sd = O.synth_state_dict(O.mbv2unet_param_shapes(10), seed=0)
    sd = O.calibrate_bn(sd, O.synth_input(8, 256, 256, seed=7))    # fixture F1; stats frozen below
    bn_keys = O.bn_stat_keys(sd)
    np.savez_compressed(os.path.join(GOLD, "mbv2unet_bnstats.npz"),
                        **{k: sd[k].numpy() for k in bn_keys})
    ref.load_state_dict(expand_aliases(sd), strict=True)
    ref.eval()
    x = O.synth_input(2, 64, 96, seed=0)
    with torch.no_grad():
        y = ref(x)
        # taps through the reference's own submodules (unet.py:34-45)
        x1 = ref.down1(x); x2 = ref.down2(x1); x3 = ref.down3(x2); x4 = ref.down4(x3); x5 = ref.down5(x4)
        u1 = ref.up1(x5, x4); u2 = ref.up2(u1, x3); u3 = ref.up3(u2, x2); u4 = ref.up4(u3, x1)
        lh = ref.outc(u4)
    am = y.argmax(1)
    print("eval logits", tuple(y.shape), "std", float(y.std()), "classes", torch.bincount(am.flatten(), minlength=10).tolist())
    top2 = y.topk(2, dim=1).values
    print("top-2 margin p0.1/median", float((top2[:, 0] - top2[:, 1]).flatten().kthvalue(max(1, am.numel() // 1000)).values),
          float((top2[:, 0] - top2[:, 1]).median()))
    taps = dict(x1=x1, x2=x2, x3=x3, x4=x4, x5=x5, u1=u1, u2=u2, u3=u3, u4=u4)
    np.savez_compressed(
        os.path.join(GOLD, "mbv2unet_eval.npz"),
        logits=y.numpy(), logits_half=lh.numpy(),
        **{f"{k}_sum": np.float64(v.double().sum().item()) for k, v in taps.items()},
        **{f"{k}_abs": np.float64(v.double().abs().sum().item()) for k, v in taps.items()},
        x5_head=x5[0, :8].numpy(), u4_head=u4[0, :4, :8, :8].numpy())

    # ---------------- MobileNetV2UNet train: reference train_model() for 2 steps ----------------
    ref.load_state_dict(expand_aliases(sd), strict=True)
    xt = O.synth_input(2, 64, 64, seed=1)
    tt = O.synth_target(2, 64, 64, seed=1)
    # (a) single fwd/bwd for loss + grads + BN stats
    ref.train()
    ref.zero_grad()
    out = ref(xt)
    loss = torch.nn.CrossEntropyLoss()(out, tt)          # main.py:99
    loss.backward()
    grads = {n: p.grad for n, p in ref.named_parameters()}
    assert grads["backbone.classifier.1.weight"] is None  # SURVEY finding 5
    sel = ["backbone.features.0.0.weight", "backbone.features.1.conv.0.0.weight", "backbone.features.2.conv.0.0.weight",
           "backbone.features.3.conv.1.0.weight", "backbone.features.18.0.weight", "backbone.features.18.1.weight",
           "backbone.features.18.1.bias", "up4.conv.conv.0.weight", "up4.conv.conv.0.bias", "up4.conv.conv.1.weight",
           "up3.conv.conv.3.weight", "outc.conv.0.weight", "outc.conv.3.weight", "outc.conv.3.bias"]
    gsum = {n: (float(g.double().sum()), float(g.double().abs().sum())) for n, g in grads.items() if g is not None}
    sd_after = ref.state_dict()
    train_blob = dict(
        loss=np.float64(loss.item()), logits=out.detach().numpy(),
        grad_names=np.array(list(gsum.keys())), grad_sum=np.array([v[0] for v in gsum.values()]),
        grad_abs=np.array([v[1] for v in gsum.values()]),
        **{"g:" + n: grads[n].numpy() for n in sel if grads[n].numel() <= 40000},
        **{"bn:" + k: sd_after[k].numpy().copy() for k in
           ["backbone.features.0.1.running_mean", "backbone.features.0.1.running_var",
            "backbone.features.18.1.running_mean", "up1.conv.conv.1.running_var", "outc.conv.1.running_mean",
            "outc.conv.1.num_batches_tracked"]})
    # (b) the reference's own training loop, 2 steps (train.py:6-79)
    ref.load_state_dict(expand_aliases(sd), strict=True)
    cwd = os.getcwd()
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td); os.makedirs("Models/obj")
        opt = torch.optim.Adam(ref.parameters(), lr=1.5e-4)  # main.py:100
        losses = []
        crit = torch.nn.CrossEntropyLoss()

        class Rec(torch.nn.Module):
            def forward(self, o, t):
                l = crit(o, t); losses.append(float(l)); return l
        loader = [(xt, tt), (O.synth_input(2, 64, 64, seed=2), O.synth_target(2, 64, 64, seed=2))]
        ref_train.train_model(ref, loader, Rec(), opt, torch.device("cpu"), epochs=1)
        ck = torch.load("Models/obj/obj_MOB_1_epoch_1.pth")
        assert len(ck) == 691
        os.chdir(cwd)
    sd2 = ref.state_dict()
    train_blob.update(
        step_losses=np.array(losses),
        **{"p2:" + n: sd2[n].numpy().copy() for n in
           ["outc.conv.3.weight", "outc.conv.3.bias", "up4.conv.conv.3.bias", "backbone.features.0.0.weight",
            "backbone.features.18.1.weight", "backbone.features.0.1.running_var"]})
    np.savez_compressed(os.path.join(GOLD, "mbv2unet_train.npz"), **train_blob)
    print("train loss", loss.item(), "loop losses", losses)

    # ---------------- plain UNet (base 16 keeps the fixture small) ----------------
    un = ref_unet.UNet(output_channels=10, base_filters=16)
    usd = O.synth_state_dict(O.unet_param_shapes(10, 16), seed=3)
    un.load_state_dict(usd, strict=True)
    xu = O.synth_input(1, 32, 48, seed=3); tu = O.synth_target(1, 32, 48, seed=3)
    un.eval()
    with torch.no_grad():
        yu = un(xu)
    un.train()
    yt = un(xu); lu = torch.nn.CrossEntropyLoss()(yt, tu); lu.backward()
    np.savez_compressed(os.path.join(GOLD, "unet_eval.npz"), logits=yu.numpy(), train_logits=yt.detach().numpy(),
                        loss=np.float64(lu.item()),
                        g_inc=un.inc.conv.conv[0].weight.grad.numpy(), g_out=un.sem_out.conv[3].weight.grad.numpy(),
                        rv=un.state_dict()["up3.conv.conv.4.running_var"].numpy())
    print("unet eval std", float(yu.std()), "loss", lu.item())
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
