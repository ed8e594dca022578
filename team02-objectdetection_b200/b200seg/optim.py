"""Fused multi-tensor Adam: ``b200seg.Adam(model.parameters(), lr=1.5e-4)`` is a drop-in for the
``optim.Adam(model.parameters(), lr=1.5e-4)`` of the reference (main.py:100; stepped at train.py:39).

Same constructor arguments, same update rule (torch/optim/adam.py, amsgrad=False, maximize=False), same ``state_dict()``
layout (``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter), so optimizer checkpoints move between the two.  The
difference is the launch count: every parameter tensor of a group is updated by ONE kernel (``b200seg_adam_multi``)
instead of ~10 ``_foreach_`` kernels over 194 tensors.  CUDA fp32 parameters only; anything else raises.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List

import torch

from . import ops
from ._cabi import check, lib, ptr


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if lr < 0.0 or eps < 0.0 or weight_decay < 0.0:
            raise ValueError("invalid Adam hyper-parameter")
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError(f"invalid betas {betas}")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        self._chunk = int(lib.b200seg_adam_chunk())
        self._plans: Dict[tuple, dict] = {}

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._plans.clear()                      # state tensors were replaced

    def _plan(self, plist: List[torch.Tensor]) -> dict:
        """Per parameter set: the chunk list (depends on the sizes only, uploaded once), the static part of the pointer
        table (p, exp_avg, exp_avg_sq, numel) and the shared step count."""
        key = tuple(id(p) for p in plist)
        plan = self._plans.get(key)
        if plan is None:
            ct, ci = [], []
            for ti, p in enumerate(plist):
                for c in range((p.numel() + self._chunk - 1) // self._chunk):
                    ct.append(ti)
                    ci.append(c)
            dev = plist[0].device
            for p in plist:
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)          # host scalar, like torch's default
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if not p.is_contiguous() or not st["exp_avg"].is_contiguous() or not st["exp_avg_sq"].is_contiguous():
                    raise RuntimeError("b200seg.Adam needs contiguous parameters and state")
            steps = {float(self.state[p]["step"]) for p in plist}
            if len(steps) != 1:
                raise RuntimeError("b200seg.Adam: parameters of one group must share the step count")
            plan = dict(n=len(ct), ct=torch.tensor(ct, dtype=torch.int32, device=dev),
                        ci=torch.tensor(ci, dtype=torch.int32, device=dev),
                        host=[torch.empty(len(plist) * 5, dtype=torch.int64).pin_memory() for _ in range(4)],
                        evs=[None] * 4, rot=0,
                        dev=torch.empty(len(plist) * 5, dtype=torch.int64, device=dev),
                        t=steps.pop(), steps=[self.state[p]["step"] for p in plist],
                        static=[(p.data_ptr(), self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr(),
                                 p.numel()) for p in plist])
            if len(self._plans) > 8:
                self._plans.clear()
            self._plans[key] = plan
        return plan

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        ops.bump_mutation_epoch()        # parameters change through raw pointers: eval-mode weight caches are stale
        for group in self.param_groups:
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            if not all(p.is_cuda for p in plist):
                raise RuntimeError("b200seg.Adam updates dense CUDA float32 parameters only (no fallback)")
            with torch.cuda.device(plist[0].device):
                self._step_group(group, plist)
        return loss

    def _step_group(self, group, plist):
        grads = []
        for p in plist:
            g = p.grad
            if not p.is_cuda or p.dtype != torch.float32 or g.dtype != torch.float32 or g.is_sparse:
                raise RuntimeError("b200seg.Adam updates dense CUDA float32 parameters only (no fallback)")
            grads.append(g if g.is_contiguous() else g.contiguous())
        plan = self._plan(plist)
        plan["t"] += 1.0
        torch._foreach_add_(plan["steps"], 1)
        t = plan["t"]
        vals = [x for (pp, mp, vp, n), g in zip(plan["static"], grads) for x in (pp, g.data_ptr(), mp, vp, n)]
        if vals != plan.get("vals"):             # gradient buffers usually come back at the same addresses
            r = plan["rot"] = (plan["rot"] + 1) % 4      # rotating pinned staging buffers: the host may run steps ahead
            if plan["evs"][r] is not None:
                plan["evs"][r].synchronize()             # upload from 4 steps ago has left this buffer
            plan["host"][r].numpy()[:] = vals
            plan["dev"].copy_(plan["host"][r], non_blocking=True)
            plan["evs"][r] = torch.cuda.Event()
            plan["evs"][r].record()
            plan["vals"] = vals
        b1, b2 = group["betas"]
        check(lib.b200seg_adam_multi(ptr(plan["dev"]), ptr(plan["ct"]), ptr(plan["ci"]), plan["n"],
                                     ctypes.c_float(group["lr"]), ctypes.c_double(1.0 - b1), ctypes.c_float(b2), ctypes.c_double(1.0 - b2),
                                     ctypes.c_float(group["eps"]), ctypes.c_float(group["weight_decay"]),
                                     ctypes.c_double(1.0 - b1 ** t), ctypes.c_double(1.0 - b2 ** t),
                                     torch.cuda.current_stream().cuda_stream), "adam_multi")
        plan["_keep"] = grads                    # contiguous copies must outlive the asynchronous kernel
