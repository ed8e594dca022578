"""T0: the drop-in modules expose exactly the reference's nn.Module surface --
constructor signatures, 691-key aliased state_dict, parameter order -- checked against key lists
frozen from the live reference (tests/golden/*_keys.json)."""
import inspect
import json
import os

import pytest
import torch

import b200seg
from oracle import unet_oracle as O
from util import GOLD, expand_aliases, fixture_sd


def _keys(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def test_mbv2unet_state_dict_layout():
    g = _keys("mbv2unet_keys.json")
    m = b200seg.MobileNetV2UNet(output_channels=10)
    sd = m.state_dict()
    assert list(sd.keys()) == [k for k, _, _ in g["keys"]]
    assert len(sd) == 691
    for k, shape, dt in g["keys"]:
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dt, k
    assert [n for n, _ in m.named_parameters()] == g["named_parameters"]
    assert [n for n, _ in m.named_buffers()] == g["named_buffers"]
    assert sum(p.numel() for p in m.parameters()) == g["n_params"] == 7830786
    assert len(list(m.parameters())) == 196
    # aliasing: downK.N.* shares storage with backbone.features.N.*  (SURVEY finding 4)
    assert len(g["alias_groups"]) == 312
    for grp in g["alias_groups"]:
        ptrs = {sd[k].data_ptr() for k in grp}
        assert len(ptrs) == 1, grp
    assert len({v.data_ptr() for v in sd.values() if v.numel() > 1}) <= 379


def test_unet_state_dict_layout():
    g = _keys("unet_keys.json")
    m = b200seg.UNet(output_channels=10)
    sd = m.state_dict()
    assert list(sd.keys()) == [k for k, _, _ in g["keys"]]
    for k, shape, dt in g["keys"]:
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dt, k
    assert [n for n, _ in m.named_parameters()] == g["named_parameters"]
    assert sum(p.numel() for p in m.parameters()) == g["n_params"] == 3364586
    assert sum(p.numel() for p in b200seg.LightUNet().parameters()) == 842977


def test_ctor_signatures_match_reference():
    assert str(inspect.signature(b200seg.MobileNetV2UNet.__init__)) == "(self, output_channels=1)"
    assert str(inspect.signature(b200seg.UNet.__init__)) == "(self, output_channels=1, base_filters=64)"
    assert str(inspect.signature(b200seg.LightUNet.__init__)) == "(self, base_filters=32)"
    m = b200seg.MobileNetV2UNet()
    for a in ("backbone", "down1", "down2", "down3", "down4", "down5", "up1", "up2", "up3", "up4", "outc",
              "final_upsample"):
        assert hasattr(m, a)
    assert m.final_upsample.align_corners is True and m.up1.up.align_corners is None
    u = b200seg.UNet()
    for a in ("inc", "down1", "down2", "down3", "up1", "up2", "up3", "sem_out"):
        assert hasattr(u, a)


def test_strict_load_of_reference_layout_checkpoint_roundtrip(tmp_path):
    sd = expand_aliases(fixture_sd())
    m = b200seg.MobileNetV2UNet(output_channels=10)
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m.down5[7][0].weight, sd["backbone.features.18.0.weight"])
    p = tmp_path / "obj_MOB_1_epoch_1.pth"          # train.py:77 naming
    torch.save(m.state_dict(), p)
    ck = torch.load(p)
    assert len(ck) == 691
    assert ck["down1.0.0.weight"].data_ptr() == ck["backbone.features.0.0.weight"].data_ptr()
    m2 = b200seg.MobileNetV2UNet(output_channels=10)
    m2.load_state_dict(ck, strict=True)
    # missing alias key must fail strict loading exactly like the reference would
    bad = dict(sd); bad.pop("down3.5.conv.2.weight")
    with pytest.raises(RuntimeError):
        b200seg.MobileNetV2UNet(output_channels=10).load_state_dict(bad, strict=True)


def test_init_statistics_follow_torchvision_and_torch_defaults():
    torch.manual_seed(0)
    m = b200seg.MobileNetV2UNet(output_channels=10)
    w = m.backbone.features[18][0].weight                    # kaiming_normal_(fan_out): std = sqrt(2/1280)
    assert abs(float(w.std()) - (2.0 / 1280) ** 0.5) < 2e-3
    bn = m.backbone.features[3].conv[1][1]
    assert torch.all(bn.weight == 1) and torch.all(bn.bias == 0) and int(bn.num_batches_tracked) == 0
    assert abs(float(m.backbone.classifier[1].weight.std()) - 0.01) < 1e-3
    assert m.up1.conv.conv[0].bias is not None and m.backbone.features[0][0].bias is None


def test_no_cpu_fallback_and_containers_refuse_to_run():
    m = b200seg.MobileNetV2UNet(output_channels=10).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="container"):
        m.up1(torch.zeros(1, 1280, 1, 1), torch.zeros(1, 64, 2, 2))
    with pytest.raises(RuntimeError, match="container"):
        m.backbone(torch.zeros(1, 3, 32, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        b200seg.CrossEntropyLoss()(torch.zeros(1, 10, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))


def test_optimizer_and_preprocess_have_no_cpu_fallback_either():
    """b200seg.Adam takes torch.optim.Adam's constructor arguments and refuses CPU parameters; preprocess_image validates
    its frame before touching the GPU."""
    p = torch.nn.Parameter(torch.zeros(4))
    opt = b200seg.Adam([p], lr=1.5e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    assert opt.defaults["lr"] == 1.5e-4 and opt.defaults["betas"] == (0.9, 0.999)
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="CUDA"):
        opt.step()
    with pytest.raises(ValueError):
        b200seg.Adam([p], lr=-1.0)
    with pytest.raises(ValueError):
        b200seg.preprocess_image(torch.zeros(8, 8, 3))                       # float frame
    with pytest.raises(ValueError):
        b200seg.preprocess_image(torch.zeros(8, 8, 1, dtype=torch.uint8))    # not 3 channels


def test_fused_schedule_replaces_the_inverted_residual_triples():
    """Eval bf16 schedule: the 16 expand-ratio-6 blocks are one step each, outconv + final upsample are one step, the 62
    convs stay covered exactly once."""
    from b200seg import engine
    m = b200seg.MobileNetV2UNet(output_channels=10)
    eng = m._get_engine()
    assert len(eng._mb_triples()) == 16
    sched = eng._schedule("bf16", "tc", 256, 512)
    fused = [s for s in sched if s.op == "mbconv"]
    assert len(fused) == 16 and len(sched) == 31
    assert sched[0].op == "stem_mb1" and [p.name for p in sched[0].parts] == [
        "backbone.features.0", "backbone.features.1.conv.0", "backbone.features.1.conv.1"]
    assert eng._schedule("bf16", "tc", 256, 512, torch.bfloat16)[0].op == "stem_mb1"  # bf16 frames too (the bench's input)
    assert eng._schedule("bf16", "tc", 258, 516, torch.bfloat16)[0].op == "stem"      # bf16 rows not 16-byte multiples: layer by layer
    assert all(s.parts[1].stride in (1, 2) and s.parts[0].src == s.src and s.parts[2].dst == s.dst for s in fused)
    assert sched[-1].op == "tail" and [p.name for p in sched[-1].parts] == ["outc.conv.0", "outc.conv.3", "final_upsample"]
    n_convs = sum(3 if s.op in ("mbconv", "stem_mb1") else 2 if s.op == "tail" else 1 for s in sched
                  if s.op in ("stem", "dw", "dense", "mbconv", "tail", "stem_mb1"))
    assert n_convs == 62
    assert eng._schedule("fp32", "simt", 256, 512) is eng.steps               # the exact path is never fused
    eng.mbconv_impl, eng.tail_impl, eng.head_impl = "unfused", "unfused", "unfused"
    assert [s.name for s in eng._schedule("bf16", "tc", 256, 512)] == [s.name for s in eng.steps]
    m17 = b200seg.MobileNetV2UNet(output_channels=17)                          # > 16 classes: the generic tail kernels
    assert m17._get_engine()._schedule("bf16", "tc", 256, 512)[-1].op == "final"


def test_schedule_covers_every_used_parameter():
    """Every parameter except the dead classifier is consumed by exactly one fused step."""
    from b200seg import engine
    m = b200seg.MobileNetV2UNet(output_channels=10)
    steps = engine.build_steps_mbv2unet(m)
    used = set()
    for s in steps:
        for mod in (s.conv, s.bn):
            if mod is not None:
                for p in mod.parameters():
                    assert id(p) not in used
                    used.add(id(p))
    names = {id(p): n for n, p in m.named_parameters()}
    unused = sorted(names[i] for i in names if i not in used)
    assert unused == ["backbone.classifier.1.bias", "backbone.classifier.1.weight"]
    assert sum(1 for s in steps if s.op in ("stem", "dw", "dense")) == 62     # 62 convs
    assert sum(1 for s in steps if s.bn is not None) == 61
    u = b200seg.UNet(10)
    us = engine.build_steps_unet(u)
    assert sum(1 for s in us if s.conv is not None) == 16 and sum(1 for s in us if s.op == "pool") == 3


def test_oracle_shape_table_is_the_distinct_tensor_list():
    shapes = O.mbv2unet_param_shapes(10)
    assert len(shapes) == 379
    m = b200seg.MobileNetV2UNet(output_channels=10)
    sd = m.state_dict()
    for k, shp, _ in shapes:
        assert tuple(sd[k].shape) == tuple(shp), k
