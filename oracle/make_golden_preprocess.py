"""Generate tests/golden/preprocess.npz by running the LIVE reference pre-processing (authoring container only).

inference.py cannot be imported (its top level loads a checkpoint and opens a camera), so the function
``preprocess_image`` (inference.py:28-46) is extracted from the reference source with ``ast`` at run time and executed
with the reference's own dependencies (cv2, torchvision.transforms) -- no reference code is copied into this repository.

Run:  python oracle/make_golden_preprocess.py        (needs /root/reference, cv2, torchvision)
"""
import ast
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/inference.py"

CASES = [  # name, source H, W, target_size (W, H) as the reference passes it
    ("down", 90, 160, (64, 32)),          # camera frame -> network size (down-scaling, the reference's use)
    ("odd", 75, 131, (96, 64)),
    ("up", 24, 40, (64, 48)),             # up-scaling exercises the clamped border rows/columns
    ("same", 32, 64, (64, 32)),           # source already at the network size
]


def load_reference_fn():
    import cv2
    import torch
    from torchvision import transforms
    tree = ast.parse(open(REF).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "preprocess_image")
    ns = {"cv2": cv2, "transforms": transforms, "device": torch.device("cpu"), "torch": torch}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), REF, "exec"), ns)
    return ns["preprocess_image"]


def main():
    fn = load_reference_fn()
    rng = np.random.default_rng(1234)
    blob = {}
    for name, h, w, ts in CASES:
        frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        # smooth half of the cases a little so that neighbouring pixels correlate like a real frame
        if name in ("down", "same"):
            frame = ((frame.astype(np.int32) + np.roll(frame, 1, 0) + np.roll(frame, 1, 1)) // 3).astype(np.uint8)
        t, img = fn(frame, target_size=ts)
        blob[f"{name}_frame"] = frame
        blob[f"{name}_target_size"] = np.asarray(ts, np.int32)
        blob[f"{name}_tensor"] = t.numpy()
        blob[f"{name}_rgb"] = np.asarray(img)
    out = os.path.join(ROOT, "tests", "golden", "preprocess.npz")
    np.savez_compressed(out, **blob)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
