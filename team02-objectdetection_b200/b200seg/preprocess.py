"""GPU version of the inference script's frame pre-processing (inference.py:28-46), batched.

    img_tensor, img = b200seg.preprocess_image(frame, target_size=(256, 128))      # same call as the reference

``frame`` is a uint8 HWC BGR image (numpy array as cv2 delivers it, or a torch uint8 tensor on the host or the GPU) or a
batch [B,H,W,3]; the result is the normalised float tensor [B,3,H,W] on the GPU and the resized RGB uint8 image
([B,H,W,3] GPU tensor; [H,W,3] for a single frame).  The resize is cv2.resize's INTER_LINEAR restated bit for bit, so the
tensor equals the reference's for every pixel.  Only the uint8 frame crosses PCIe: 3 bytes per source pixel.
"""
from __future__ import annotations

import ctypes

import torch

from ._cabi import BF16, F32, check, lib, ptr

MEAN = (0.485, 0.456, 0.406)      # inference.py:37
STD = (0.229, 0.224, 0.225)       # inference.py:38


def preprocess_image(image, target_size=(256, 128), device=None, dtype=torch.float32, mean=MEAN, std=STD, out=None,
                     want_rgb: bool = True):
    if not torch.is_tensor(image):
        import numpy as np
        image = torch.from_numpy(np.ascontiguousarray(image))
    if image.dtype != torch.uint8 or image.dim() not in (3, 4) or image.shape[-1] != 3:
        raise ValueError(f"expected a uint8 HWC (or BHWC) BGR frame, got {tuple(image.shape)} {image.dtype}")
    single = image.dim() == 3
    frames = image.unsqueeze(0) if single else image
    if not frames.is_cuda:
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        frames = frames.to(dev, non_blocking=True)
    frames = frames.contiguous()
    if dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("preprocess_image produces float32 or bfloat16")
    B, Hs, Ws, _ = frames.shape
    W, H = int(target_size[0]), int(target_size[1])             # cv2 order: (width, height)
    if out is None:
        out = torch.empty((B, 3, H, W), device=frames.device, dtype=dtype)
    rgb = torch.empty((B, H, W, 3), device=frames.device, dtype=torch.uint8) if want_rgb else None
    f = ctypes.c_float
    with torch.cuda.device(frames.device):           # launch on the frames' device, whatever the current one is
        check(lib.b200seg_preprocess_u8(ptr(frames), B, Hs, Ws, ptr(out), BF16 if out.dtype == torch.bfloat16 else F32, ptr(rgb),
                                        H, W, f(mean[0]), f(mean[1]), f(mean[2]), f(std[0]), f(std[1]), f(std[2]),
                                        torch.cuda.current_stream().cuda_stream), "preprocess_u8")
    return out, (None if rgb is None else rgb[0] if single else rgb)
