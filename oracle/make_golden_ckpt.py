"""A checkpoint WRITTEN BY THE REFERENCE ITSELF, frozen as a fixture -- authoring container only; TEST INFRASTRUCTURE.

The reference's only persistent interface is the ``.pth`` its training loop writes (``train.py:77``:
``torch.save(model.state_dict(), 'Models/obj/obj_MOB_1_epoch_N.pth')``) and its consumers read back
(``inference.py:24``, ``convert.py:23``).  This script runs the reference's own ``train_model()`` for one epoch of two
synthetic batches on its own ``UNet(output_channels=10, base_filters=16)`` (the plain UNet keeps the file at ~0.5 MB; a
MobileNetV2UNet checkpoint is 27 MB) and copies the file it wrote -- byte for byte -- to
``tests/golden/ref_unet16_epoch_1.pth``, next to the reference's eval logits for that checkpoint on a seeded frame
(``tests/golden/ref_unet16_epoch_1.npz``).  tests/test_checkpoint.py loads it into the drop-in with strict=True.

Run:  python oracle/make_golden_ckpt.py        (needs /root/reference)
"""
import os
import shutil
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import unet_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    sys.path.insert(0, "/root/reference")
    from src.unet import UNet            # the reference's own module (unet.py:124-147)
    from src.train import train_model    # the reference's own loop (train.py:6-79)
    torch.manual_seed(0)
    model = UNet(output_channels=10, base_filters=16)
    loader = [(O.synth_input(2, 32, 48, seed=s), O.synth_target(2, 32, 48, seed=s)) for s in (11, 12)]
    opt = torch.optim.Adam(model.parameters(), lr=1.5e-4)          # main.py:100
    crit = torch.nn.CrossEntropyLoss()                            # main.py:99
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)
        os.makedirs("Models/obj")
        train_model(model, loader, crit, opt, torch.device("cpu"), epochs=1)
        os.chdir(cwd)
        shutil.copyfile(os.path.join(td, "Models/obj/obj_MOB_1_epoch_1.pth"), os.path.join(GOLD, "ref_unet16_epoch_1.pth"))
    # what the reference's consumers do with it (inference.py:23-25): fresh model, load_state_dict, eval
    m2 = UNet(output_channels=10, base_filters=16)
    m2.load_state_dict(torch.load(os.path.join(GOLD, "ref_unet16_epoch_1.pth"), map_location="cpu"))
    m2.eval()
    x = O.synth_input(1, 32, 48, seed=13)
    with torch.no_grad():
        y = m2(x)
    np.savez_compressed(os.path.join(GOLD, "ref_unet16_epoch_1.npz"), x=x.numpy(), logits=y.numpy())
    print("wrote", os.path.getsize(os.path.join(GOLD, "ref_unet16_epoch_1.pth")), "bytes; logits", tuple(y.shape), float(y.abs().max()))


if __name__ == "__main__":
    main()
