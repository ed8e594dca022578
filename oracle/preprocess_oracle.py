"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's frame pre-processing (inference.py:28-46):

    cv2.resize(image, target_size)  ->  cv2.cvtColor(BGR2RGB)  ->  transforms.ToTensor()  ->  transforms.Normalize(mean, std)
    -> unsqueeze(0)

The arithmetic lives in two un-vendored dependencies (requirements.txt, unpinned; installed here: opencv-python 4.13.0,
torchvision 0.26.0):
  * cv2.resize(INTER_LINEAR) on uint8 is OpenCV's fixed-point bilinear (modules/imgproc/src/resize.cpp: resizeGeneric_ with
    HResizeLinear / VResizeLinear<uchar,int,short>): source coordinate fx = (float)((dx + 0.5) * scale - 0.5), 11-bit
    coefficients cvRound(c * 2048), horizontal pass in int32, vertical pass
        dst = ((b0 * (S0 >> 4) >> 16) + (b1 * (S1 >> 4) >> 16) + 2) >> 2.
    Horizontally a clamped tap zeroes its fraction; vertically the two source rows are clamped individually and the
    fractions are kept (that asymmetry is OpenCV's and is needed for bit-exactness on up-scaled borders).
  * ToTensor: uint8 HWC -> float32 CHW / 255;  Normalize: (t - mean) / std in float32.
Pinned: tests/golden/preprocess.npz holds outputs of the reference's own preprocess_image() run in the authoring
container (oracle/make_golden.py), and tests/test_oracle_golden.py checks this file against them bit for bit.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np

MEAN = (0.485, 0.456, 0.406)      # inference.py:37
STD = (0.229, 0.224, 0.225)       # inference.py:38


def _coeffs(ssize: int, dsize: int, zero_clamped: bool):
    scale = ssize / dsize
    idx = np.zeros(dsize, np.int64)
    a = np.zeros((dsize, 2), np.int64)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if zero_clamped:
            if s < 0:
                f, s = np.float32(0.0), 0
            if s >= ssize - 1:
                f, s = np.float32(0.0), ssize - 1
        idx[d] = s
        a[d, 0] = int(np.rint(np.float64(np.float32(1.0) - f) * 2048))      # cvRound: round half to even
        a[d, 1] = int(np.rint(np.float64(f) * 2048))
    return idx, a


def resize_linear_u8(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    """cv2.resize(src, (dw, dh)) for uint8 HWC, interpolation=INTER_LINEAR (bit-exact restatement)."""
    sh, sw = src.shape[:2]
    xi, xa = _coeffs(sw, dw, True)
    yi, ya = _coeffs(sh, dh, False)
    s = src.astype(np.int64)
    x1 = np.minimum(xi + 1, sw - 1)
    hor = s[:, xi, :] * xa[None, :, 0, None] + s[:, x1, :] * xa[None, :, 1, None]
    y0, y1 = np.clip(yi, 0, sh - 1), np.clip(yi + 1, 0, sh - 1)
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (hor[y0] >> 4)) >> 16) + ((b1 * (hor[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def preprocess_image(image: np.ndarray, target_size=(256, 128)):
    """inference.py:28-46.  image: uint8 HWC BGR.  Returns (float32 [1,3,H,W], uint8 RGB [H,W,3])."""
    img = resize_linear_u8(image, int(target_size[0]), int(target_size[1]))
    img = img[:, :, ::-1].copy()                                        # BGR -> RGB
    t = img.astype(np.float32).transpose(2, 0, 1) / np.float32(255.0)   # ToTensor
    mean = np.asarray(MEAN, np.float32)[:, None, None]
    std = np.asarray(STD, np.float32)[:, None, None]
    t = (t - mean) / std                                                # Normalize
    return t[None], img
