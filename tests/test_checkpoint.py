"""checkpoint fidelity with a file the REFERENCE wrote, and the export graph convert.py needs.

tests/golden/ref_unet16_epoch_1.pth is the byte-for-byte output of the reference's own ``train_model()`` (train.py:77) on its
own ``UNet(10, base_filters=16)``; ``ref_unet16_epoch_1.npz`` holds the reference's eval logits for that checkpoint
(oracle/make_golden_ckpt.py)."""
import os

import numpy as np
import pytest
import torch

import b200seg
from oracle import unet_oracle as O
from util import GOLD, expand_aliases, fixture_sd, rel_err

CKPT = os.path.join(GOLD, "ref_unet16_epoch_1.pth")


def _gold():
    g = np.load(os.path.join(GOLD, "ref_unet16_epoch_1.npz"))
    return torch.from_numpy(g["x"]), torch.from_numpy(g["logits"])


def test_reference_written_checkpoint_loads_strict_and_the_oracle_reproduces_its_logits():
    ck = torch.load(CKPT, map_location="cpu")             # inference.py:24 / convert.py:23
    m = b200seg.UNet(output_channels=10, base_filters=16)
    assert list(ck.keys()) == list(m.state_dict().keys())                       # same keys, same order
    res = m.load_state_dict(ck, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in m.state_dict().items():
        assert v.dtype == ck[k].dtype and torch.equal(v, ck[k]), k
    x, ref = _gold()
    y = O.unet_forward({k: v for k, v in ck.items()}, x)                        # the oracle, pinned to a real checkpoint
    assert rel_err(y, ref) < 2e-5


def test_saving_the_drop_in_gives_a_file_the_reference_layout_reads_back(tmp_path):
    ck = torch.load(CKPT, map_location="cpu")
    m = b200seg.UNet(output_channels=10, base_filters=16)
    m.load_state_dict(ck, strict=True)
    p = tmp_path / "obj_MOB_1_epoch_2.pth"
    torch.save(m.state_dict(), p)                                               # train.py:77
    back = torch.load(p, map_location="cpu")
    assert list(back.keys()) == list(ck.keys())
    assert all(torch.equal(back[k], ck[k]) for k in ck)


def test_export_graph_matches_the_reference_logits_and_traces():
    """convert.py:26-42 traces the model; the traceable graph of the drop-in must be the reference's function."""
    ck = torch.load(CKPT, map_location="cpu")
    m = b200seg.UNet(output_channels=10, base_filters=16)
    m.load_state_dict(ck, strict=True)
    m.eval()
    g = b200seg.torch_graph(m)
    x, ref = _gold()
    with torch.no_grad():
        y = g(x)
    assert rel_err(y, ref) < 2e-5
    assert g.model.inc.conv.conv[0].weight is m.inc.conv.conv[0].weight        # shared parameters, not copies
    traced = torch.jit.trace(g, x)                                             # what torch.onnx.export does first
    with torch.no_grad():
        assert rel_err(traced(torch.cat([x, x.flip(0)])), torch.cat([ref, ref.flip(0)])) < 2e-5   # dynamic batch axis


def test_export_graph_of_mobilenetv2_unet_equals_the_frozen_reference_logits():
    sd = fixture_sd()
    m = b200seg.MobileNetV2UNet(output_channels=10)
    m.load_state_dict(expand_aliases(sd), strict=True)
    m.eval()
    x = O.synth_input(2, 64, 96, seed=0)                      # the input the frozen reference logits were produced on
    with torch.no_grad():
        y = b200seg.torch_graph(m)(x)
    frozen = torch.from_numpy(np.load(os.path.join(GOLD, "mbv2unet_eval.npz"))["logits"])
    assert rel_err(y, frozen) < 2e-5
    assert rel_err(y, O.mobilenetv2_unet_forward(sd, x)) < 2e-5


def test_onnx_export_runs_when_the_exporter_is_installed(tmp_path):
    pytest.importorskip("onnx")
    m = b200seg.UNet(output_channels=10, base_filters=16)
    m.load_state_dict(torch.load(CKPT, map_location="cpu"), strict=True)
    m.eval()
    x, _ = _gold()
    p = tmp_path / "unet16.onnx"
    torch.onnx.export(b200seg.torch_graph(m), x, str(p), export_params=True, opset_version=12, do_constant_folding=True,
                      input_names=["input"], output_names=["output"],
                      dynamic_axes={"input": {0: "batch_size"}, "output": {0: "batch_size"}})      # convert.py:28-42
    assert p.stat().st_size > 100_000


def test_class_map_lut_equals_the_reference_loop():
    """BDD100KDataset.py:23-35,66-69: mapped = zeros_like(mask); for s, t in class_map.items(): mapped[mask == s] = t."""
    class_map = {0: 1, 13: 2, 6: 3, 7: 4, 11: 5, 1: 6, 14: 7, 15: 8, 17: 9, 18: 9, 12: 9}
    rng = np.random.default_rng(0)
    mask = rng.integers(0, 256, size=(3, 37, 53), dtype=np.uint8)
    mapped = np.zeros_like(mask)
    for s, t in class_map.items():
        mapped[mask == s] = t
    lut = b200seg.class_map_lut(class_map)
    assert lut.dtype == torch.uint8 and lut.numel() == 256
    assert np.array_equal(lut.numpy()[mask], mapped)
    with pytest.raises(ValueError):
        b200seg.class_map_lut({300: 1})
    with pytest.raises(RuntimeError, match="CUDA"):
        b200seg.remap_labels(torch.from_numpy(mask), lut)                      # no CPU fallback
