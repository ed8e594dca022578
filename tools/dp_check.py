"""T4 on 2 GPUs: a 2-rank data-parallel step equals the single-process computation that runs
each shard through its own BatchNorm statistics and averages the gradients; replicas stay bit-identical.
Launch: python tools/dp_check.py   (spawns 2 ranks itself)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "team02-objectdetection_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def worker(rank, world, port, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import b200seg
    from b200seg import dp
    from oracle import unet_oracle as O
    from util import expand_aliases, fixture_sd
    sd = fixture_sd()
    m = b200seg.MobileNetV2UNet(output_channels=10)
    if rank == 0:
        m.load_state_dict(expand_aliases(sd), strict=True)     # rank 1 keeps its random init until the broadcast
    m = m.to(dev)
    dp.broadcast_model(m)
    dp.attach(m, bucket_bytes=2 << 20)
    m.train()
    x, t = O.synth_input(4, 64, 64, seed=1), O.synth_target(4, 64, 64, seed=1)
    xs, ts = x[rank * 2:(rank + 1) * 2].to(dev), t[rank * 2:(rank + 1) * 2].to(dev)
    loss = b200seg.CrossEntropyLoss()(m(xs), ts)
    loss.backward()
    grads = {n: p.grad.detach().cpu() for n, p in m.named_parameters() if p.grad is not None}
    q.put((rank, float(loss), grads))
    dist.barrier()
    dist.destroy_process_group()


def main():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, 2, 29577, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
    # single-process expectation: each shard separately on GPU 0 (per-replica BN), gradients averaged
    import b200seg
    from oracle import unet_oracle as O
    from util import expand_aliases, fixture_sd
    sd = fixture_sd()
    x, t = O.synth_input(4, 64, 64, seed=1), O.synth_target(4, 64, 64, seed=1)
    dev = torch.device("cuda", 0)
    exp = None
    for r in range(2):
        m = b200seg.MobileNetV2UNet(output_channels=10)
        m.load_state_dict(expand_aliases(sd), strict=True)
        m = m.to(dev).train()
        b200seg.CrossEntropyLoss()(m(x[r * 2:(r + 1) * 2].to(dev)), t[r * 2:(r + 1) * 2].to(dev)).backward()
        gr = {n: p.grad.detach().cpu() for n, p in m.named_parameters() if p.grad is not None}
        exp = gr if exp is None else {k: (exp[k] + gr[k]) for k in gr}
    exp = {k: v / 2 for k, v in exp.items()}
    g0, g1 = res[0][2], res[1][2]
    assert set(g0) == set(exp) and len(g0) == 194
    same = all(torch.equal(g0[k], g1[k]) for k in g0)
    worst = max(float((g0[k] - exp[k]).abs().max() / exp[k].abs().max().clamp_min(1e-12)) for k in g0 if float(exp[k].abs().max()) > 1e-6)
    print(f"ranks bit-identical: {same}; worst rel deviation from the shard-wise expectation: {worst:.2e}; losses {res[0][1]:.6f} {res[1][1]:.6f}")
    assert same and worst < 1e-3
    print("DP CHECK OK")


if __name__ == "__main__":
    main()
