"""Training forward/backward of the drop-in models (model.train(): train.py:24,36,38).

One ``torch.autograd.Function`` spans the whole network: ``forward`` runs the fused step schedule with
train-mode BatchNorm (batch statistics, running-stat update, ``num_batches_tracked += 1``) and keeps, per
layer, the pre-BN conv output ``z``, the BN statistics and the activation output; ``backward`` walks the
schedule in reverse and returns one gradient per used parameter, so ``loss.backward()`` /
``optimizer.step()`` at the reference call sites work unchanged.  ``backbone.classifier`` is not on the path:
its ``.grad`` stays ``None`` exactly as in the reference (SURVEY finding 5).

Kernels: forward convs and the dense data gradients (same conv with transposed/flipped weights) go through
``conv_simt`` (f32) or ``conv_tc`` (bf16), dense weight gradients through ``conv_wgrad`` / ``conv_wgrad_tc``;
everything else is in csrc/train_ops.cu.  Gradient accumulation for tensors with two consumers (skip
connections, block inputs with a shortcut) rides on the conv kernels' ``+residual`` epilogue or the ``acc``
argument of the adjoint kernels -- there is no separate add pass.

Steady state: after ``engine.graph_after`` steps with the same input shape, the forward and the backward are
each captured into a CUDA graph over static buffers (weight re-packing included, so in-place optimizer updates
are picked up); a step then costs two graph launches of host time instead of ~1600 kernel launches.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops
from .ops import ACT_NONE

TRACE = None      # set to a list to collect (phase, step name, start event, end event) per schedule step (eager only)


def _tick():
    if TRACE is None:
        return None
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def _train_params(engine) -> List[torch.nn.Parameter]:
    ps = []
    for s in engine.steps:
        if s.conv is not None:
            ps.append(s.conv.weight)
            if s.conv.bias is not None:
                ps.append(s.conv.bias)
        if s.bn is not None:
            ps += [s.bn.weight, s.bn.bias]
    return ps


# ----------------------------------------------------------------------------------------------------
# weight repack: all layers in one launch (b200seg_pack_weights_multi) into persistent operand buffers
# ----------------------------------------------------------------------------------------------------
def _train_packs(engine, tc: bool):
    """Persistent packed-weight buffers of the training step + the device table that fills them.  Returns
    {step name: dict(wp=..., wk=..., wt=...)} after launching the repack for the CURRENT parameter values.
    (before, every layer cost 3-6 tiny permute/cast/flip kernels per step, ~280 launches.)"""
    import ctypes
    from ._cabi import check, lib, ptr
    convs = [s for s in engine.steps if s.op in ("stem", "dw", "dense")]
    dev = convs[0].conv.weight.device
    key = (tc, dev, tuple(s.conv.weight.data_ptr() for s in convs))
    plan = getattr(engine, "_train_pack_plan", None)
    if plan is None or plan["key"] != key:
        chunk = int(lib.b200seg_pack_chunk())
        rows, views, ct, ci = [], {}, [], []
        n16 = n32 = 0
        metas = []
        for s in convs:
            w = s.conv.weight
            if not w.is_contiguous() or w.dtype != torch.float32:
                raise TypeError("training keeps contiguous fp32 master weights")
            cout, cin, k = w.shape[0], w.shape[1], w.shape[2]
            kk = k * k
            if s.op == "stem":
                kind, cp, nf, nd = 2, cout, kk * cin * cout, 0
            elif s.op == "dw":
                kind, cp, nf, nd = 3, cout, kk * cout, kk * cout
            else:
                kind = 0 if tc else 1
                cp = max(cout, s.pad_cout) if s.pad_cout else cout
                nf = nd = cp * kk * cin
            metas.append((s, kind, cout, cin, kk, cp, nf, nd))
            if kind == 0:
                n16 += (nf + 7) // 8 * 8 + (nd + 7) // 8 * 8
            else:
                n32 += (nf + 3) // 4 * 4 + (nd + 3) // 4 * 4
        buf16 = torch.zeros(max(n16, 8), device=dev, dtype=torch.bfloat16)
        buf32 = torch.zeros(max(n32, 4), device=dev, dtype=torch.float32)
        o16 = o32 = 0
        for ti, (s, kind, cout, cin, kk, cp, nf, nd) in enumerate(metas):
            if kind == 0:
                fwd = buf16[o16:o16 + nf]; o16 += (nf + 7) // 8 * 8
                dg = buf16[o16:o16 + nd]; o16 += (nd + 7) // 8 * 8
            else:
                fwd = buf32[o32:o32 + nf]; o32 += (nf + 3) // 4 * 4
                dg = buf32[o32:o32 + nd] if nd else None; o32 += (nd + 3) // 4 * 4
            if s.op == "stem":
                views[s.name] = dict(wp=fwd.view(s.conv.weight.shape[2], s.conv.weight.shape[3], cin, cout))
            elif s.op == "dw":
                views[s.name] = dict(wp=fwd.view(kk, cout), wpf=dg.view(kk, cout))
            else:
                views[s.name] = dict(wk=fwd.view(cp, kk * cin), wt=dg.view(cin, kk * cp))
            # struct PackEntry {w, fwd, dgrad (8 bytes each); cout, cin, kk, cout_pad, kind, pad (4 bytes each)}
            rows += [s.conv.weight.data_ptr(), fwd.data_ptr(), dg.data_ptr() if dg is not None else 0,
                     (cin << 32) | cout, (cp << 32) | kk, kind]
            n = cp * cin * kk
            for c in range((n + chunk - 1) // chunk):
                ct.append(ti); ci.append(c)
        plan = dict(key=key, views=views, buf16=buf16, buf32=buf32, n=len(ct),
                    table=torch.tensor(rows, dtype=torch.int64, device=dev),
                    ct=torch.tensor(ct, dtype=torch.int32, device=dev), ci=torch.tensor(ci, dtype=torch.int32, device=dev))
        engine._train_pack_plan = plan
    check(lib.b200seg_pack_weights_multi(ptr(plan["table"]), ptr(plan["ct"]), ptr(plan["ci"]), plan["n"],
                                         torch.cuda.current_stream().cuda_stream), "pack_weights_multi")
    return plan["views"]


# ----------------------------------------------------------------------------------------------------
# the two passes as plain functions over tensors (no autograd): used eagerly and under graph capture
# ----------------------------------------------------------------------------------------------------
def run_forward(engine, x: torch.Tensor, mode: str):
    sdt = torch.bfloat16 if mode == "bf16" else torch.float32
    tc = mode == "bf16" and (engine.dense_impl or "tc") == "tc"
    env: Dict[str, torch.Tensor] = {"x": x}
    saved: Dict[str, dict] = {}
    ops.zero_pool.reset()
    ops.zero_pool32.reset()
    packs = _train_packs(engine, tc)       # every layer's operand layouts from the current fp32 weights, one launch
    counters = []                          # BatchNorm.num_batches_tracked, bumped with one fused launch at the end
    for s in engine.steps:
        rec: dict = {}
        _e0 = _tick()
        if s.op in ("stem", "dw", "dense"):
            cout = s.conv.weight.shape[0]
            bias = s.conv.bias.detach().float() if s.conv.bias is not None else None
            src = env[s.src]
            pk = packs[s.name]
            if s.op == "stem":
                rec["wp"] = pk["wp"]
                z = ops.conv3x3_smallcin(src, rec["wp"], bias, s.stride, ACT_NONE, sdt)
            elif s.op == "dw":
                rec["wp"], rec["wpf"] = pk["wp"], pk["wpf"]
                z = ops.dwconv3x3(src, rec["wp"], bias, s.stride, ACT_NONE)
            else:
                if s.pad_cout and cout < s.pad_cout and bias is not None:
                    bias = torch.cat([bias, bias.new_zeros(s.pad_cout - cout)], 0)
                rec["wk"], rec["wt"] = pk["wk"], pk["wt"]     # [Cout_pad, taps*Cin] and its transposed/tap-flipped twin
                if tc:
                    z = ops.conv_tc(src, rec["wk"], bias, s.taps, ACT_NONE, None, flags=engine.tc_flags)
                else:
                    z = ops.conv_simt(src, rec["wk"], bias, s.taps, ACT_NONE, None)
            rec["z"] = z
            if s.bn is not None:
                res = env[s.res] if (s.op == "dense" and s.res) else None
                a, sv = ops.bn_train_forward(z, s.bn.weight.detach().float(), s.bn.bias.detach().float(),
                                             s.bn.running_mean, s.bn.running_var, s.bn.eps,
                                             s.bn.momentum if s.bn.momentum is not None else 0.1, s.act, res)
                counters.append(s.bn.num_batches_tracked)
                rec["sv"] = sv
                env[s.dst] = a
            else:
                env[s.dst] = z
        elif s.op == "upcat":
            env[s.dst] = ops.upsample2x_concat(env[s.res], env[s.src])
        elif s.op == "pool":
            env[s.dst] = ops.maxpool2x2(env[s.src])
        elif s.op == "final":
            env[s.dst] = ops.upsample2x_ac_nchw(env[s.src], engine.out_ch, torch.float32)
        elif s.op == "to_nchw":
            env[s.dst] = ops.nhwc_to_nchw(env[s.src], engine.out_ch, torch.float32)
        else:  # pragma: no cover
            raise AssertionError(s.op)
        saved[s.name] = rec
        if _e0 is not None:
            TRACE.append(("fwd", s.name, _e0, _tick()))
    if counters:
        torch._foreach_add_(counters, 1)
    return env, saved


def run_backward(engine, env, saved, mode: str, dout: torch.Tensor, emit) -> None:
    """Walk the schedule in reverse; every parameter gradient is handed to ``emit(param, grad)`` as soon as it
    exists (the data-parallel reducer starts a bucket's all-reduce from there)."""
    sdt = torch.bfloat16 if mode == "bf16" else torch.float32
    tc = mode == "bf16" and (engine.dense_impl or "tc") == "tc"
    g: Dict[str, torch.Tensor] = {"out": dout}
    ops.zero_pool.reset()
    ops.zero_pool32.reset()
    for s in reversed(engine.steps):
        rec = saved[s.name]
        _e0 = _tick()
        if s.op == "final":
            g[s.src] = ops.final_bwd(g.pop(s.dst), sdt)
        elif s.op == "to_nchw":
            g[s.src] = ops.nchw_to_nhwc_pad(g.pop(s.dst), env[s.src].shape[-1], sdt)
        elif s.op == "upcat":
            dcat = g.pop(s.dst)
            cs = env[s.res].shape[-1]
            dskip, dx = ops.upcat_bwd(dcat, cs, g.get(s.res))
            g[s.res] = dskip
            assert s.src not in g
            g[s.src] = dx
        elif s.op == "pool":
            g[s.src] = ops.maxpool_bwd(env[s.src], g.pop(s.dst), g.get(s.src))
        else:
            da = g.pop(s.dst)
            z = rec["z"]
            if s.bn is not None:
                if s.op == "dense" and s.res:           # shortcut: the same gradient flows to the block input
                    assert s.res not in g
                    g[s.res] = da
                dz, dgamma, dbeta = ops.bn_train_backward(da, z, rec["sv"], s.act)
                emit(s.bn.weight, dgamma)
                emit(s.bn.bias, dbeta)
            else:
                dz = da                                   # conv + bias only (the last 1x1 of outconv)
            w = s.conv.weight
            cout = w.shape[0]
            if s.conv.bias is not None:
                emit(s.conv.bias, ops.colsum(dz)[:cout].contiguous())
            src = env[s.src]
            if s.op == "stem":
                dwp = ops.smallcin_wgrad(src, dz, s.stride)            # [3,3,Cin,Cout]
                emit(w, dwp.permute(3, 2, 0, 1).contiguous())
            elif s.op == "dw":
                dw9 = ops.dw_wgrad(src, dz, s.stride)                  # [9,C]
                emit(w, dw9.t().reshape(cout, 1, 3, 3).contiguous())
                if s.stride == 1 and g.get(s.src) is None:
                    # stride 1: the data gradient IS the forward depthwise conv of dz with the taps flipped -> the
                    # register-blocked forward kernel (2.5x faster than the generic gather kernel)
                    g[s.src] = ops.dwconv3x3(dz, rec["wpf"], None, 1, ACT_NONE)
                else:
                    g[s.src] = ops.dw_dgrad(dz, rec["wp"], tuple(src.shape), s.stride, g.get(s.src))
            else:
                cin = src.shape[-1]
                k = 3 if s.taps == 9 else 1
                dwk = ops.conv_wgrad_tc(src, dz, s.taps) if tc else ops.conv_wgrad(src, dz, s.taps)   # [Cout_pad, taps*Cin]
                emit(w, dwk[:cout].reshape(cout, k, k, cin).permute(0, 3, 1, 2).contiguous())
                # dgrad = the same conv with W transposed (and the 3x3 taps flipped): packed with the forward operand
                wt = rec["wt"]                                          # [Cin, taps*Cout_pad]
                if tc:
                    g[s.src] = ops.conv_tc(dz, wt, None, s.taps, ACT_NONE, g.get(s.src), flags=engine.tc_flags)
                else:
                    g[s.src] = ops.conv_simt(dz, wt, None, s.taps, ACT_NONE, g.get(s.src))
        if _e0 is not None:
            TRACE.append(("bwd", s.name, _e0, _tick()))


# ----------------------------------------------------------------------------------------------------
# CUDA-graph replay of the two passes
# ----------------------------------------------------------------------------------------------------
class _StepGraph:
    """Forward and backward graphs of one (input shape, dtype, mode), sharing one memory pool."""

    def __init__(self, engine, x: torch.Tensor, mode: str):
        self.x = torch.empty_like(x)
        self.x.copy_(x)
        # torch.cuda.graph() does not run the captured work, but BatchNorm's num_batches_tracked bump is captured
        # like any other kernel, so nothing is double counted.
        # build the persistent repack plan (allocations, table upload) OUTSIDE the capture
        _train_packs(engine, mode == "bf16" and (engine.dense_impl or "tc") == "tc")
        torch.cuda.synchronize()
        self.fwd = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.fwd):
            self.env, self.saved = run_forward(engine, self.x, mode)
        self.out = self.env["out"]
        self.dout = torch.zeros_like(self.out)
        self.grads: Dict[int, torch.Tensor] = {}
        self.bwd = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.bwd, pool=self.fwd.pool()):
            run_backward(engine, self.env, self.saved, mode, self.dout, lambda p, gr: self.grads.__setitem__(id(p), gr))
        self.pending = False      # a forward has been replayed whose backward has not run yet


class _TrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, allow_graph, x, *params):
        mode = engine._mode(x)
        x = x.contiguous()
        ctx.engine, ctx.mode, ctx.graph = engine, mode, None
        sg = _graph_for(engine, x, mode) if allow_graph else None
        if sg is not None:
            sg.x.copy_(x)
            sg.fwd.replay()
            sg.pending = True
            ctx.graph = sg
            out = sg.out.clone()
        else:
            with torch.no_grad():
                ctx.env, ctx.saved = run_forward(engine, x, mode)
            out = ctx.env["out"]
        return out.to(x.dtype) if x.dtype != torch.float32 else out

    @staticmethod
    def backward(ctx, dout):
        engine = ctx.engine
        red = getattr(engine, "reducer", None)      # data parallel: bucketed all-reduce (overlapped with backward when eager)
        if red is not None:
            red.reset()
        params = _train_params(engine)
        sg = ctx.graph
        if sg is not None:
            sg.dout.copy_(dout)
            sg.bwd.replay()
            sg.pending = False
            if red is not None:
                for p in red.params:                 # buckets fill in backward order; all-reduces start as they fill
                    red.add(p, sg.grads[id(p)])
                pgrad = red.finish()
                return (None, None, None, *[pgrad[id(p)].clone().to(p.dtype) for p in params])
            return (None, None, None, *[sg.grads[id(p)].clone().to(p.dtype) for p in params])
        pgrad: Dict[int, torch.Tensor] = {}

        def emit(p, gr):
            pgrad[id(p)] = gr
            if red is not None:
                red.add(p, gr)
        with torch.no_grad():
            run_backward(engine, ctx.env, ctx.saved, ctx.mode, dout.float().contiguous(), emit)
        if red is not None:
            pgrad = red.finish()                     # rank-averaged views into the flat buckets
        ctx.env = ctx.saved = None
        return (None, None, None, *[(pgrad[id(p)].to(p.dtype) if id(p) in pgrad else None) for p in params])


def _graph_for(engine, x, mode) -> Optional[_StepGraph]:
    """The step graph to replay for this call, or None (eager): graphs need grad mode, no tracing, a warmed-up
    shape, and no forward still waiting for its backward (two forwards would share the static buffers)."""
    if not engine.use_graphs or TRACE is not None or torch.cuda.is_current_stream_capturing():
        return None
    key = ("train", tuple(x.shape), x.dtype, mode, engine.dense_impl, engine.tc_flags, x.device)
    ent = engine._graphs.get(key)
    if ent is None:
        if len(engine._graphs) >= 3:
            engine._graphs.clear()
        ent = engine._graphs[key] = {"seen": 0, "graph": None}
    ent["seen"] += 1
    if ent["graph"] is None and ent["seen"] > engine.graph_after:
        ent["graph"] = _StepGraph(engine, x, mode)
    sg = ent["graph"]
    if sg is None or sg.pending:
        return None
    return sg


def forward_train(engine, x: torch.Tensor) -> torch.Tensor:
    params = _train_params(engine)
    engine._check_input(x)
    if params[0].dtype != torch.float32:
        raise TypeError("training keeps fp32 master weights: use model.float() and engine.precision='bf16' "
                        "(or torch.autocast) for bf16 activations")
    if not torch.is_grad_enabled():
        # model.train() under no_grad (e.g. BN calibration): forward only, still updates running stats
        return _TrainFn.apply(engine, False, x, *[p.detach() for p in params])
    return _TrainFn.apply(engine, True, x, *params)
