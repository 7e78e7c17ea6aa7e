"""Multi-GPU path on ONE GPU: the halo kernels against torch index ops, and every rank's share of a node-partitioned /
batch-sharded layer call executed back to back on cuda:0 with the exchange replaced by the gathers it is equivalent to.
Forward rows must be bit-identical to the unpartitioned call (same messages, same order); gradients within tolerance.
The real 2..8-rank NCCL run is tests/dist_check.py (torchrun; `gpurun --gpus N`)."""
import ctypes as C

import numpy as np
import pytest
import torch

import ngpde
from ngpde import _lib, distributed as D, ops, partition as P, workloads
from common import product_fwd_bwd, relerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


@pytest.mark.parametrize("d", [1, 3, 64])
def test_rows_gather_and_segment_add_kernels(d):
    rng = np.random.default_rng(d)
    n, m = 500, 1300
    x = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32)).to(DEV)
    rows = torch.from_numpy(rng.integers(0, n, m).astype(np.int32)).to(DEV)
    lib = _lib.load()
    out = torch.empty((m, d), device=DEV)
    _lib.check(lib.ngpde_rows_gather(x.data_ptr(), rows.data_ptr(), m, d, out.data_ptr(), ops._stream(x.device)))
    assert torch.equal(out, x[rows.long()])
    # segment add: rows visited in ascending receive position, compared with a sequential float32 loop
    order = np.argsort(rows.cpu().numpy(), kind="stable")
    rs = rows.cpu().numpy()[order]
    head = np.concatenate([[True], rs[1:] != rs[:-1]])
    seg_rows, seg_ptr = rs[head], np.concatenate([np.nonzero(head)[0], [m]])
    dst = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32))
    src = torch.from_numpy(rng.standard_normal((m, d)).astype(np.float32))
    ref = dst.clone()
    for k in order:  # stable order == ascending position within a row
        r = int(rows[k])
        ref[r] = ref[r] + src[k]
    dd, sd = dst.to(DEV), src.to(DEV)
    i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(DEV)
    a, b, c = i32(seg_rows), i32(seg_ptr), i32(order)
    _lib.check(lib.ngpde_rows_segment_add(dd.data_ptr(), sd.data_ptr(), a.data_ptr(), b.data_ptr(), c.data_ptr(),
                                          len(seg_rows), d, ops._stream(dd.device)))
    assert torch.equal(dd.cpu(), ref)


def test_rows_put_writes_per_peer_buffers():
    rng = np.random.default_rng(0)
    n, d = 300, 8
    x = torch.from_numpy(rng.standard_normal((n, d)).astype(np.float32)).to(DEV)
    counts = [40, 0, 25]
    rows = torch.from_numpy(rng.integers(0, n, sum(counts)).astype(np.int32)).to(DEV)
    bufs = [torch.zeros((max(c, 1) + 2, d), device=DEV) for c in counts]
    offs = [1, 0, 2]  # row offsets inside the peers' buffers
    peer_dst = torch.tensor([b.data_ptr() + 4 * d * o for b, o in zip(bufs, offs)], dtype=torch.int64, device=DEV)
    peer_ptr = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int64, device=DEV)
    _lib.check(_lib.load().ngpde_rows_put(x.data_ptr(), rows.data_ptr(), peer_ptr.data_ptr(), peer_dst.data_ptr(), 3,
                                          sum(counts), d, ops._stream(x.device)))
    o = 0
    for b, c, off in zip(bufs, counts, offs):
        assert torch.equal(b[off:off + c], x[rows[o:o + c].long()])
        o += c


def _simulate_partitioned(w, world, dy):
    """All ranks' work on one GPU; the exchange is replaced by index_select of the full arrays (what it delivers)."""
    x_full = w.x.detach()
    N = w.n_nodes
    y = torch.empty((dy.shape[0], N), device=DEV)
    dx = torch.zeros((x_full.shape[0], N), device=DEV, dtype=torch.float64)
    dps = None
    for r in range(world):
        pl = D.PartitionedLayer(w.layer, w.graph, r, world, DEV)
        p = pl.part
        l2g = torch.from_numpy(p.local_to_global()).to(DEV)
        x_local = x_full[:, l2g].detach().clone().requires_grad_(True)
        ca = ngpde.ComponentArray(w.ps)
        ca.data.requires_grad_(True)
        y_local, _ = w.layer(x_local, ca, pl.local_state(w.st))
        y[:, p.lo:p.hi] = y_local[:, :p.n_owned].detach()
        dy_local = torch.zeros_like(y_local)
        dy_local[:, :p.n_owned] = dy[:, p.lo:p.hi]
        y_local.backward(dy_local)
        dx.index_add_(1, l2g, x_local.grad.double())  # what the reverse exchange + segment add accumulates
        dps = ca.data.grad.double() if dps is None else dps + ca.data.grad.double()
    return y, dx.float(), dps.float()


@pytest.mark.parametrize("name,kw,world", [("c3", {"side": 24}, 4), ("c4", {"n_nodes": 3000, "chs": 8, "hidden": 16}, 3),
                                            ("c1", {}, 2)])
def test_node_partitioned_layer_matches_single_gpu(name, kw, world):
    w = workloads.WORKLOADS[name](DEV, **kw)
    g = torch.Generator().manual_seed(0)
    y0, _ = w.layer(w.x, w.ps, w.st)
    dy = torch.randn(tuple(y0.shape), generator=g).to(DEV)
    y0, dx0, dp0 = product_fwd_bwd(w.layer, w.x, w.ps, w.st, dy)
    y, dx, dp = _simulate_partitioned(w, world, dy)
    assert torch.equal(y, y0), "owned rows must be bit-identical to the single-GPU forward"
    assert relerr(dx, dx0) <= TOL and relerr(dp, dp0) <= TOL, (relerr(dx, dx0), relerr(dp, dp0))


def test_ensemble_shards_match_the_full_batch():
    w = workloads.c2_mppde(DEV, n_per=32, n_graphs=6, hidden=24)
    g = torch.Generator().manual_seed(0)
    y0, _ = w.layer(w.x, w.ps, w.st)
    dy = torch.randn(tuple(y0.shape), generator=g).to(DEV)
    y0, dx0, dp0 = product_fwd_bwd(w.layer, w.x, w.ps, w.st, dy)
    world = 4
    dps = 0
    for r in range(world):
        gl, sh = D.shard_ensemble(w.graph, r, world, DEV)
        st = ngpde.updategraph(w.st, gl)
        sl = slice(sh.node_lo, sh.node_hi)
        y, dx, dp = product_fwd_bwd(w.layer, w.x[:, sl], w.ps, st, dy[:, sl])
        assert torch.equal(y, y0[:, sl]) and relerr(dx, dx0[:, sl]) <= TOL
        dps = dps + dp.double()
    assert relerr(dps.float(), dp0) <= TOL
