"""SURVEY 8f rank 3/4 on the GPU: a reference-written checkpoint through the CUDA path, the label remap kernel, and the
double-buffered input feed that replaces train.py:32-33."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import b200seg  # noqa: E402
from util import GOLD, rel_err  # noqa: E402

DEV = "cuda"


def test_reference_written_checkpoint_through_the_cuda_path():
    ck = torch.load(os.path.join(GOLD, "ref_unet16_epoch_1.pth"), map_location="cpu")        # inference.py:24
    g = np.load(os.path.join(GOLD, "ref_unet16_epoch_1.npz"))
    x, ref = torch.from_numpy(g["x"]), torch.from_numpy(g["logits"])
    m = b200seg.UNet(output_channels=10, base_filters=16)
    m.load_state_dict(ck, strict=True)
    m = m.to(DEV).eval()
    with torch.no_grad():
        y = m(x.to(DEV)).cpu()
    assert rel_err(y, ref) < 1e-4                                   # north_star: fp32 logits within 1e-4 relative
    assert (y.argmax(1) == ref.argmax(1)).float().mean().item() >= 0.999
    with torch.no_grad():
        mask = m.predict_mask(x.to(DEV)).cpu()
    assert (mask.long() == ref.argmax(1)).float().mean().item() >= 0.999


@pytest.mark.parametrize("shape", [(2, 128, 256), (1, 37, 53), (3, 5), (4096 * 3 + 17,)])
def test_remap_labels_equals_the_reference_loop(shape):
    class_map = {0: 1, 13: 2, 6: 3, 7: 4, 11: 5, 1: 6, 14: 7, 15: 8, 17: 9, 18: 9, 12: 9}    # BDD100KDataset.py:23-35
    rng = np.random.default_rng(1)
    mask = rng.integers(0, 20, size=shape, dtype=np.uint8)
    mapped = np.zeros_like(mask)
    for s, t in class_map.items():                                                           # BDD100KDataset.py:66-69
        mapped[mask == s] = t
    got = b200seg.remap_labels(torch.from_numpy(mask).to(DEV), b200seg.class_map_lut(class_map))
    assert got.dtype == torch.int64 and tuple(got.shape) == shape
    assert np.array_equal(got.cpu().numpy(), mapped.astype(np.int64))                        # .long(), BDD100KDataset.py:75


def test_device_feeder_yields_the_loader_batches_in_order():
    g = torch.Generator().manual_seed(0)
    lut = b200seg.class_map_lut({0: 1, 13: 2})
    batches = []
    for i in range(5):
        b = 4 if i < 4 else 2                                       # ragged last batch, as a DataLoader without drop_last
        batches.append((torch.randn(b, 3, 32, 64, generator=g), torch.randint(0, 20, (b, 32, 64), generator=g, dtype=torch.uint8)))
    feeder = b200seg.DeviceFeeder(batches, DEV, lut=lut)
    assert len(feeder) == 5
    seen = 0
    for (x, y), (xh, yh) in zip(feeder, batches):
        assert x.is_cuda and y.is_cuda and y.dtype == torch.int64
        z = x * 2.0                                                 # work enqueued on the batch while the next one uploads
        assert torch.equal(x.cpu(), xh) and torch.equal(z.cpu(), xh * 2.0)
        assert torch.equal(y.cpu(), lut.long()[yh.long()])
        seen += 1
    assert seen == 5
    # int64 targets pass through untouched; the loop is reusable (second epoch)
    plain = [(torch.randn(2, 3, 8, 8, generator=g), torch.randint(0, 10, (2, 8, 8), generator=g)) for _ in range(3)]
    f2 = b200seg.DeviceFeeder(plain, DEV)
    for epoch in range(2):
        for (x, y), (xh, yh) in zip(f2, plain):
            assert torch.equal(x.cpu(), xh) and torch.equal(y.cpu(), yh)


def test_device_feeder_drives_a_training_loop_like_train_py():
    """train.py:31-42 with the loop header swapped for the feeder: the losses of uploading each batch synchronously."""
    from oracle import unet_oracle as O
    torch.manual_seed(0)
    batches = [(O.synth_input(2, 32, 64, seed=s), O.synth_target(2, 32, 64, seed=s)) for s in range(3)]

    def run(feed):
        torch.manual_seed(1)
        m = b200seg.UNet(output_channels=10, base_filters=16).to(DEV).train()
        opt = b200seg.Adam(m.parameters(), lr=1.5e-4)
        crit = b200seg.CrossEntropyLoss()
        out = []
        for x, y in feed:
            x, y = x.to(DEV), y.to(DEV)
            opt.zero_grad()
            loss = crit(m(x), y)
            loss.backward()
            opt.step()
            out.append(loss.item())
        return out
    a, b = run(b200seg.DeviceFeeder(batches, DEV)), run(batches)
    assert len(a) == len(b) == 3
    assert max(abs(u - v) for u, v in zip(a, b)) < 1e-5        # atomics reorder the f32 loss sum from run to run
