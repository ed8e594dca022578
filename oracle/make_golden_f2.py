"""Fixture F2: train the LIVE reference (/root/reference, MobileNetV2UNet(10)) on the synthetic road-scene
generator, freeze the trained network and the reference's own outputs on it.  Authoring container only; TEST
INFRASTRUCTURE, never imported by the product.

  tests/golden/f2_weights.npz   trained weights of every tensor on the forward path (backbone.classifier dropped: never
                                used, SURVEY finding 5), conv weights int8 per output channel (oracle.f2_state_dict)
  tests/golden/f2_eval.npz      the reference's eval logits on the quantised network: full logits at 2x64x96, checksums +
                                a strided sample + argmax histogram at 2x256x512 (the shape the bf16 parity test uses),
                                the reference's OWN bf16 floor at both shapes, and its train-mode loss/gradient
                                checksums at 4x64x128 (the shape of the bf16 training test)

Run:  python oracle/make_golden_f2.py [steps]     (default 600 Adam(1e-3) steps, B=8, 128x256: ~6 min on 8 vCPU)
"""
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402
from oracle.make_golden import load_reference, expand_aliases  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def strip(sd):
    return {k: v for k, v in sd.items() if k.startswith(("backbone.features.", "up", "outc."))}


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    torch.set_num_threads(os.cpu_count() or 8)
    ref_unet, _ = load_reference()
    torch.manual_seed(0)
    ref = ref_unet.MobileNetV2UNet(output_channels=10)        # torchvision's own random init (unet.py:12 with weights=None)
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    crit = torch.nn.CrossEntropyLoss()
    ref.train()
    t0 = time.time()
    for it in range(steps):
        x, t = O.road_scene_batch(8, 128, 256, seed=it)
        opt.zero_grad()
        loss = crit(ref(x), t)
        loss.backward()
        opt.step()
        if it % 25 == 0 or it == steps - 1:
            print(f"step {it:4d} loss {loss.item():.4f}  ({time.time() - t0:.0f} s)", flush=True)
    final_loss = float(loss)

    # freeze: int8-per-channel weights; the reference is then evaluated on EXACTLY the network every box reconstructs
    blob = O.f2_quantize(strip(ref.state_dict()))
    np.savez_compressed(os.path.join(GOLD, "f2_weights.npz"), **blob)
    sd = O.f2_state_dict(np.load(os.path.join(GOLD, "f2_weights.npz")))
    full = dict(ref.state_dict())
    full.update(expand_aliases(sd))
    ref.load_state_dict(full, strict=True)
    ref.eval()

    out = {}
    meta = {"steps": steps, "final_train_loss": final_loss}
    for tag, (b, h, w, seed) in {"small": (2, 64, 96, 9001), "full": (2, 256, 512, 9002)}.items():
        x, t = O.road_scene_batch(b, h, w, seed=seed)
        with torch.no_grad():
            y = ref(x)
            y_or = O.mobilenetv2_unet_forward(sd, x)
        assert float((y - y_or).abs().max()) <= 2e-5 * float(y.abs().max()), "oracle drifted from the live reference"
        am = y.argmax(1)
        acc = float((am == t).float().mean())
        top2 = y.topk(2, dim=1).values
        margin = (top2[:, 0] - top2[:, 1]).flatten()
        # the reference's own bf16 run (model.bfloat16() on the same weights/input): the noise floor of ANY bf16 path
        rb = ref_unet.MobileNetV2UNet(output_channels=10)
        rb.load_state_dict(full, strict=True)
        rb = rb.bfloat16().eval()
        with torch.no_grad():
            yb = rb(x.bfloat16()).float()
        rng = float(y.max() - y.min())
        floor = dict(max_rel=float((yb - y).abs().max() / y.abs().max()), mean_over_range=float((yb - y).abs().mean() / rng),
                     max_over_range=float((yb - y).abs().max() / rng), argmax_agree=float((yb.argmax(1) == am).float().mean()))
        meta[tag] = dict(shape=[b, h, w], seed=seed, pixel_accuracy=acc, logits_std=float(y.std()),
                         margin_p001=float(margin.kthvalue(max(1, margin.numel() // 1000)).values),
                         margin_median=float(margin.median()), reference_bf16_floor=floor)
        print(tag, json.dumps(meta[tag]))
        if tag == "small":
            out["small_logits"] = y.numpy()
        else:
            out["full_sum"] = np.float64(y.double().sum().item())
            out["full_abs"] = np.float64(y.double().abs().sum().item())
            out["full_sample"] = y[:, :, ::8, ::8].numpy().copy()
            out["full_argmax_hist"] = torch.bincount(am.flatten(), minlength=10).numpy()
            out["full_argmax_sample"] = am[:, ::4, ::4].to(torch.uint8).numpy().copy()

    # train-mode reference numbers at the shape of the bf16 training test
    ref.load_state_dict(full, strict=True)
    ref.train()
    ref.zero_grad()
    x, t = O.road_scene_batch(4, 64, 128, seed=9100)
    loss = crit(ref(x), t)
    loss.backward()
    names = [n for n, p in ref.named_parameters() if p.grad is not None]
    out["train_loss"] = np.float64(loss.item())
    out["train_grad_names"] = np.array(names)
    out["train_grad_abs"] = np.array([float(dict(ref.named_parameters())[n].grad.double().abs().sum()) for n in names])
    for n in ("outc.conv.3.weight", "up4.conv.conv.3.weight", "up4.conv.conv.0.weight", "up3.conv.conv.3.weight"):
        out["g:" + n] = dict(ref.named_parameters())[n].grad.numpy().copy()
    meta["train"] = dict(shape=[4, 64, 128], seed=9100, loss=float(loss))
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(GOLD, "f2_eval.npz"), **out)
    for f in ("f2_weights.npz", "f2_eval.npz"):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
