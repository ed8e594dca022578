"""T4 on 2 GPUs: data-parallel training over NCCL.

  1. one 2-rank step equals the single-process computation that runs each shard through its own BatchNorm statistics
     and averages the gradients (eager path);
  2. five 2-rank steps with the fused Adam: steps 3..5 replay the captured forward/backward graphs with the per-bucket
     all-reduces INSIDE the backward graph.  Replicas must stay bit-identical (parameters, gradients, BN buffers) and the
     trajectory must agree with the same five steps run eagerly (use_graphs=False) within Adam's run-to-run noise.
Launch: python tools/dp_check.py   (spawns 2 ranks itself)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

LR, STEPS = 1.5e-4, 5


def make_model(dev, load):
    import b200seg
    from util import expand_aliases, fixture_sd
    m = b200seg.MobileNetV2UNet(output_channels=10)
    if load:
        m.load_state_dict(expand_aliases(fixture_sd()), strict=True)
    return m.to(dev)


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import b200seg
    from b200seg import dp
    from oracle import unet_oracle as O
    x, t = O.synth_input(4, 64, 64, seed=1), O.synth_target(4, 64, 64, seed=1)
    xs, ts = x[rank * 2:(rank + 1) * 2].to(dev), t[rank * 2:(rank + 1) * 2].to(dev)
    crit = b200seg.CrossEntropyLoss()

    # ---- 1. single step, eager
    m = make_model(dev, load=rank == 0)          # rank 1 keeps its random init until the broadcast
    dp.broadcast_model(m)
    dp.attach(m, bucket_bytes=2 << 20)
    m.train()
    loss = crit(m(xs), ts)
    loss.backward()
    grads = {n: p.grad.detach().cpu() for n, p in m.named_parameters() if p.grad is not None}
    out = {"rank": rank, "loss1": float(loss), "grads1": grads}

    # ---- 2. five steps, graphs on / off
    for tag, use_graphs in (("graph", True), ("eager", False)):
        m = make_model(dev, load=rank == 0)
        dp.broadcast_model(m)
        dp.attach(m, bucket_bytes=2 << 20)
        eng = m._get_engine()
        eng.use_graphs, eng.graph_after = use_graphs, 2
        opt = b200seg.Adam(m.parameters(), lr=LR)
        m.train()
        losses = []
        for _ in range(STEPS):
            opt.zero_grad()
            loss = crit(m(xs), ts)
            loss.backward()
            opt.step()
            losses.append(float(loss))
        replayed = any(e.get("graph") is not None for e in eng._graphs.values())
        out[tag] = dict(losses=losses, replayed=replayed,
                        params={n: p.detach().cpu() for n, p in m.named_parameters()},
                        grads={n: p.grad.detach().cpu() for n, p in m.named_parameters() if p.grad is not None},
                        bufs={n: b.detach().cpu() for n, b in m.named_buffers()})
    q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def main():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, 29577, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda r: r["rank"])
    for p in procs:
        p.join(timeout=120)
    # single-process expectation of step 1: each shard separately on GPU 0 (per-replica BN), gradients averaged
    import b200seg
    from oracle import unet_oracle as O
    x, t = O.synth_input(4, 64, 64, seed=1), O.synth_target(4, 64, 64, seed=1)
    dev = torch.device("cuda", 0)
    exp = None
    for r in range(2):
        m = make_model(dev, True).train()
        b200seg.CrossEntropyLoss()(m(x[r * 2:(r + 1) * 2].to(dev)), t[r * 2:(r + 1) * 2].to(dev)).backward()
        gr = {n: p.grad.detach().cpu().clone() for n, p in m.named_parameters() if p.grad is not None}
        exp = gr if exp is None else {k: (exp[k] + gr[k]) for k in gr}
    exp = {k: v / 2 for k, v in exp.items()}
    g0, g1 = res[0]["grads1"], res[1]["grads1"]
    assert set(g0) == set(exp) and len(g0) == 194
    same = all(torch.equal(g0[k], g1[k]) for k in g0)
    worst = max(float((g0[k] - exp[k]).abs().max() / exp[k].abs().max().clamp_min(1e-12)) for k in g0 if float(exp[k].abs().max()) > 1e-6)
    print(f"step 1: ranks bit-identical: {same}; worst rel deviation from the shard-wise expectation: {worst:.2e}; "
          f"losses {res[0]['loss1']:.6f} {res[1]['loss1']:.6f}")
    assert same and worst < 1e-3

    a, b = res[0]["graph"], res[1]["graph"]
    assert a["replayed"] and b["replayed"], "the graph path was not taken"
    same_p = all(torch.equal(a["params"][k], b["params"][k]) for k in a["params"])
    same_g = all(torch.equal(a["grads"][k], b["grads"][k]) for k in a["grads"])
    same_b = all(torch.equal(a["bufs"][k], b["bufs"][k]) for k in a["bufs"] if "running" not in k)   # BN stats are per replica
    e = res[0]["eager"]
    dmax = max(float((a["params"][k] - e["params"][k]).abs().max()) for k in a["params"])
    dmean = sum(float((a["params"][k] - e["params"][k]).abs().sum()) for k in a["params"]) / sum(v.numel() for v in a["params"].values())
    dloss = max(abs(u - v) for u, v in zip(a["losses"], e["losses"]))
    print(f"{STEPS} steps (3..{STEPS} replayed, all-reduce inside the backward graph): replicas bit-identical params {same_p} grads {same_g} "
          f"counters {same_b}; graph vs eager trajectory: max |dp| {dmax / LR:.2f} lr, mean {dmean / LR:.4f} lr, max |dloss| {dloss:.2e}; "
          f"losses {['%.5f' % v for v in a['losses']]}")
    assert same_p and same_g and same_b
    assert dmax <= 2.05 * LR * STEPS and dmean < 0.1 * LR and dloss < 2e-3
    print("DP CHECK OK")


if __name__ == "__main__":
    main()
