"""GPU frame pre-processing (b200seg_preprocess_u8 behind b200seg.preprocess_image) against the oracle restatement of
inference.py:28-46 and the reference outputs frozen in tests/golden/preprocess.npz.  Integer work (the resized uint8 image)
must be bit-exact; the float32 tensor is produced by the same three IEEE operations and must be bit-exact too."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import b200seg  # noqa: E402
from oracle import preprocess_oracle as P  # noqa: E402
from util import gold  # noqa: E402


def test_preprocess_matches_frozen_reference_outputs():
    g = gold("preprocess.npz")
    for name in ("down", "odd", "up", "same"):
        ts = tuple(int(v) for v in g[f"{name}_target_size"])
        t, img = b200seg.preprocess_image(g[f"{name}_frame"], target_size=ts)
        assert t.shape == (1, 3, ts[1], ts[0]) and t.dtype == torch.float32 and t.is_cuda
        assert np.array_equal(img.cpu().numpy(), g[f"{name}_rgb"]), name
        assert np.array_equal(t.cpu().numpy(), g[f"{name}_tensor"]), name


@pytest.mark.parametrize("hs,ws,tw,th", [(720, 1280, 512, 256), (1080, 1920, 256, 128), (256, 512, 512, 256), (97, 131, 512, 256),
                                         (480, 640, 1280, 736), (3, 5, 16, 8), (1, 1, 4, 4)])
def test_preprocess_matches_oracle_on_camera_sizes(hs, ws, tw, th):
    rng = np.random.default_rng(hs * 7 + ws)
    frames = rng.integers(0, 256, (3, hs, ws, 3), dtype=np.uint8)
    t, rgb = b200seg.preprocess_image(torch.from_numpy(frames), target_size=(tw, th))
    assert t.shape == (3, 3, th, tw) and rgb.shape == (3, th, tw, 3)
    for b in range(3):
        tr, ir = P.preprocess_image(frames[b], (tw, th))
        assert np.array_equal(rgb[b].cpu().numpy(), ir)
        assert np.array_equal(t[b].cpu().numpy(), tr[0])


def test_preprocess_bf16_output_feeds_the_model():
    """bf16 output = the float32 result rounded once; the tensor goes straight into the drop-in model."""
    rng = np.random.default_rng(5)
    frames = torch.from_numpy(rng.integers(0, 256, (2, 90, 160, 3), dtype=np.uint8)).pin_memory()
    t32, _ = b200seg.preprocess_image(frames, target_size=(128, 64))
    t16, _ = b200seg.preprocess_image(frames, target_size=(128, 64), dtype=torch.bfloat16)
    assert torch.equal(t16, t32.bfloat16())
    m = b200seg.MobileNetV2UNet(output_channels=10).cuda().bfloat16().eval()
    with torch.no_grad():
        mask = m.predict_mask(t16)
    assert mask.shape == (2, 64, 128) and mask.dtype == torch.uint8


def test_preprocess_same_size_fast_path_bf16_and_no_rgb():
    rng = np.random.default_rng(9)
    frames = rng.integers(0, 256, (2, 64, 128, 3), dtype=np.uint8)
    t16, rgb = b200seg.preprocess_image(torch.from_numpy(frames), target_size=(128, 64), dtype=torch.bfloat16, want_rgb=False)
    assert rgb is None
    for b in range(2):
        tr, _ = P.preprocess_image(frames[b], (128, 64))
        assert torch.equal(t16[b].cpu(), torch.from_numpy(tr[0]).bfloat16())


def test_preprocess_rejects_bad_input():
    with pytest.raises(ValueError):
        b200seg.preprocess_image(torch.zeros(4, 4, 3))                       # not uint8
    with pytest.raises(ValueError):
        b200seg.preprocess_image(torch.zeros(4, 4, 4, dtype=torch.uint8))    # not 3 channels
