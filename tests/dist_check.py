"""Real multi-rank check of the node-partitioned path (run under torchrun on >= 2 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/dist_check.py [--mode nccl|put|native] [--order morton]

Every rank builds the same synthetic GNOConv / VMHConv workload, runs its share through PartitionedLayer (halo exchange
over NCCL or direct peer stores), and rank 0 compares the gathered result with the unpartitioned single-GPU call:
forward rows bit-identical, dx and the all-reduced parameter gradient within 1e-5."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="nccl", choices=["nccl", "put", "native"])
    ap.add_argument("--order", default=None, choices=[None, "morton"])
    args = ap.parse_args()
    import ngpde
    from ngpde import distributed as D, workloads
    from common import product_fwd_bwd, relerr
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for name, kw in (("c4", {"n_nodes": 20000, "chs": 16, "hidden": 32}), ("c3", {"side": 64})):
        w = workloads.WORKLOADS[name](dev, **kw)
        gen = torch.Generator().manual_seed(0)
        y_full, _ = w.layer(w.x, w.ps, w.st)
        dy = torch.randn(tuple(y_full.shape), generator=gen).to(dev)
        y0, dx0, dp0 = product_fwd_bwd(w.layer, w.x, w.ps, w.st, dy)
        pl = D.PartitionedLayer(w.layer, w.graph, rank, world, dev, mode=args.mode, order=args.order)
        p = pl.part
        own = pl.owned_global.to(dev)
        x_owned = pl.owned(w.x).detach().clone().requires_grad_(True)
        ca = ngpde.ComponentArray(w.ps)
        ca.data.requires_grad_(True)
        for it in range(2):  # twice: buffers of the put mode are reused across calls
            x_owned.grad = None
            ca.data.grad = None
            y, _ = pl(x_owned, ca, pl.local_state(w.st))
            y.backward(dy[:, own])
            if args.mode == "native":
                pl.exchange.comm.allreduce_sum(ca.data.grad)  # NCCL through the C ABI's own communicator
            else:
                D.allreduce_gradients([ca.data.grad])
        e_y = float((y.detach() != y0[:, own]).sum().item())
        e_dx = relerr(x_owned.grad, dx0[:, own])
        e_dp = relerr(ca.data.grad, dp0)
        stats = torch.tensor([e_y, e_dx, e_dp], dtype=torch.float64, device=dev)
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{w.name} world={world} mode={args.mode} order={args.order}: halo rows {p.n_halo} of {p.n_owned} owned; "
                  f"forward mismatches {int(stats[0])}, dx rel err {stats[1]:.2e}, dps rel err {stats[2]:.2e}", flush=True)
        ok = ok and stats[0].item() == 0 and stats[1].item() <= 1e-5 and stats[2].item() <= 1e-5
    # peer-memory one-shot all-reduce vs NCCL: same sum (fixed rank order; NCCL's order may differ in the last bit)
    n = 25666
    par = D.PeerAllReduce(n, dev)
    gen = torch.Generator().manual_seed(100 + rank)
    v = torch.randn(n, generator=gen).to(dev)
    for it in range(3):
        par.buffer.copy_(v * (it + 1))
        got = par.reduce().clone()
        ref = (v * (it + 1)).clone()
        dist.all_reduce(ref)
        err = float((got - ref).abs().max().item() / ref.abs().max().item())
        same = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(same, got)
        identical = all(torch.equal(same[0], t) for t in same)
        if rank == 0:
            print(f"peer all-reduce it {it}: rel err vs NCCL {err:.2e}, bit-identical on all ranks: {identical}", flush=True)
        ok = ok and err <= 1e-6 and identical
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit("dist_check FAILED")
    if rank == 0:
        print("dist_check ok", flush=True)


if __name__ == "__main__":
    main()
