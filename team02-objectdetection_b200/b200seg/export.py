"""Export graph: the reference's ``convert.py:23-42`` loads a checkpoint into the model and hands it to
``torch.onnx.export`` -- a TRACER, which needs a forward made of standard ATen ops.  The drop-in's forward is a chain of
hand-written CUDA kernels behind a C ABI and cannot be traced, so

    graph = b200seg.torch_graph(model)               # nn.Module, same parameters (shared, not copied), any device
    torch.onnx.export(graph, dummy_input, path, opset_version=12, ...)        # convert.py:28-42 unchanged otherwise

returns the model's static schedule (``engine.build_steps_*``) spelled in ``torch.nn.functional`` ops, NCHW, exactly the
op sequence of ``src/unet.py`` + torchvision's ``mobilenet_v2``: conv2d / batch_norm / relu / hardtanh(0, 6) /
interpolate(bilinear) / cat / max_pool2d.  This is an EXPORT artefact: ``model.forward`` never routes through it (there is
no CPU or eager fallback on the hot path), and tests pin it to the frozen reference logits.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn
import torch.nn.functional as F

from .ops import ACT_NONE, ACT_RELU, ACT_RELU6


class TorchGraph(nn.Module):
    """The schedule of a ``b200seg.MobileNetV2UNet`` / ``UNet`` as standard PyTorch ops over the SAME parameter tensors."""

    def __init__(self, model):
        super().__init__()
        self.model = model                      # registers the parameters / buffers (state_dict keys: "model.<reference key>")
        self._steps = model._get_engine().steps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        env: Dict[str, torch.Tensor] = {"x": x}
        training = self.model.training
        for s in self._steps:
            if s.op in ("stem", "dw", "dense"):
                c = s.conv
                y = F.conv2d(env[s.src], c.weight, c.bias, c.stride, c.padding, c.dilation, c.groups)
                if s.bn is not None:
                    bn = s.bn
                    y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, training,
                                     bn.momentum if bn.momentum is not None else 0.0, bn.eps)
                if s.act == ACT_RELU:
                    y = F.relu(y)
                elif s.act == ACT_RELU6:
                    y = F.hardtanh(y, 0.0, 6.0)
                elif s.act != ACT_NONE:  # pragma: no cover
                    raise AssertionError(s.act)
                if s.op == "dense" and s.res:
                    y = env[s.res] + y              # InvertedResidual shortcut (tv:mobilenetv2.py:61-62)
                env[s.dst] = y
            elif s.op == "upcat":                   # unet.py:97,100-103
                up = F.interpolate(env[s.src], scale_factor=2, mode="bilinear", align_corners=False)
                env[s.dst] = torch.cat([env[s.res], up], dim=1)
            elif s.op == "pool":                    # unet.py:85
                env[s.dst] = F.max_pool2d(env[s.src], 2)
            elif s.op == "final":                   # unet.py:30,49
                env[s.dst] = F.interpolate(env[s.src], scale_factor=2, mode="bilinear", align_corners=True)
            elif s.op == "to_nchw":
                env[s.dst] = env[s.src]
            else:  # pragma: no cover
                raise AssertionError(s.op)
        return env["out"]


def torch_graph(model) -> TorchGraph:
    """Traceable pure-PyTorch graph of ``model`` (for ``torch.onnx.export`` / ``torch.jit.trace``, convert.py:26-42)."""
    if not hasattr(model, "_get_engine"):
        raise TypeError("torch_graph expects a b200seg model")
    g = TorchGraph(model)
    g.train(model.training)
    return g
