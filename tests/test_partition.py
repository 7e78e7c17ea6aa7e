"""Host logic of the multi-GPU path: node partition / halo lists / batch sharding, and the exchange plumbing
over a 2-rank gloo group on CPU.  The device primitives (pack, segment add) are CUDA-only in the product; the gloo test
substitutes torch index ops for exactly those two calls so that counts, ordering and the collective wiring are checked."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ngpde
from ngpde import partition as P
from ngpde import workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def brute_force_partition(s, t, bounds, rank):
    """Plain-Python restatement used as the checker for partition_nodes."""
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    eids = [k for k in range(len(s)) if lo <= t[k] < hi]
    halo = sorted({int(s[k]) for k in eids if not (lo <= s[k] < hi)})
    pos = {g: (hi - lo) + i for i, g in enumerate(halo)}
    s_loc = [int(s[k]) - lo if lo <= s[k] < hi else pos[int(s[k])] for k in eids]
    t_loc = [int(t[k]) - lo for k in eids]
    return eids, halo, s_loc, t_loc


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("kind", ["grid", "random"])
def test_partition_lists_match_brute_force(world, kind):
    rng = np.random.default_rng(3)
    if kind == "grid":
        s, t, _ = workloads.grid_edges(9, 7, 8, rng)
        n = 63
    else:
        n = 40
        s, t = rng.integers(0, n, 300), rng.integers(0, n, 300)
    parts = [P.partition_nodes(s, t, n, world, r) for r in range(world)]
    b = parts[0].bounds
    assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
    assert sum(p.edge_ids.size for p in parts) == len(s)           # every edge has exactly one owner
    for r, p in enumerate(parts):
        eids, halo, s_loc, t_loc = brute_force_partition(s, t, b, r)
        assert p.edge_ids.tolist() == eids                          # original relative order is kept
        assert p.halo_global.tolist() == halo
        assert p.s_local.tolist() == s_loc and p.t_local.tolist() == t_loc
        l2g = p.local_to_global()
        assert np.array_equal(l2g[p.s_local], s[p.edge_ids]) and np.array_equal(l2g[p.t_local], t[p.edge_ids])
        # what I receive from q is what q sends to me, in the same order
        off = 0
        for q in range(world):
            c = int(p.recv_counts[q])
            mine = p.halo_global[off:off + c]
            qs = parts[q]
            so = int(qs.send_counts[:r].sum())
            assert np.array_equal(mine, qs.send_local[so:so + int(qs.send_counts[r])] + qs.lo)
            assert int(qs.peer_recv_offset[r]) == off or c == 0
            off += c
        # segment lists: every send slot appears once, grouped by owned row, ascending position inside a group
        assert sorted(p.seg_pos.tolist()) == list(range(int(p.send_counts.sum())))
        for u in range(len(p.seg_rows)):
            q = p.seg_pos[p.seg_ptr[u]:p.seg_ptr[u + 1]]
            assert np.all(p.send_local[q] == p.seg_rows[u]) and np.all(np.diff(q) > 0)
        assert len(set(p.seg_rows.tolist())) == len(p.seg_rows)


def test_balanced_bounds_balance_edges():
    rng = np.random.default_rng(0)
    s, t, _ = workloads.radius_edges(4000, 16.0, rng)
    for world in (2, 4, 8):
        parts = [P.partition_nodes(s, t, 4000, world, r) for r in range(world)]
        st = P.partition_summary(parts)
        assert st["edges_max_over_mean"] < 1.15, st
        # strips of a spatially sorted radius graph import only a thin halo
        assert st["halo_frac_max"] < 0.6, st


def test_shard_batch_whole_graphs_only():
    rng = np.random.default_rng(1)
    s, t = workloads.path_edges(16, 6, rng)
    got = []
    for r in range(4):
        sh = P.shard_batch(s, t, 96, 6, 4, r)
        assert sh.node_lo == sh.g0 * 16 and sh.node_hi == sh.g1 * 16
        assert np.array_equal(sh.s_local + sh.node_lo, s[sh.edge_ids])
        assert np.all((sh.t_local >= 0) & (sh.t_local < sh.node_hi - sh.node_lo))
        got.append(sh.edge_ids)
    assert np.array_equal(np.sort(np.concatenate(got)), np.arange(len(s)))
    with pytest.raises(ValueError):  # an edge between graphs of different ranks is not a batch
        P.shard_batch(np.array([0, 20]), np.array([20, 0]), 96, 6, 4, 0)
    with pytest.raises(ValueError):
        P.shard_batch(s, t, 95, 6, 4, 0)


# ---------------------------------------------------------------------------------------------------------------------
# world_size-2 gloo run of HaloExchange / allreduce_gradients
# ---------------------------------------------------------------------------------------------------------------------

def _free_port():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import ngpde  # noqa: F401
    from ngpde import distributed as D
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class CpuExchange(D.HaloExchange):  # test stand-ins for the two CUDA-only device primitives
            def _pack(self, x, rows):
                return x.index_select(0, rows.long())

            def _segment_add(self, dst, src):
                for u in range(self.seg_rows.numel()):
                    for q in range(int(self.seg_ptr[u]), int(self.seg_ptr[u + 1])):
                        dst[int(self.seg_rows[u])] += src[int(self.seg_pos[q])]

        rng = np.random.default_rng(5)
        n = 60
        s, t = rng.integers(0, n, 400), rng.integers(0, n, 400)
        part = P.partition_nodes(s, t, n, world, rank)
        ex = CpuExchange(part, "cpu")
        d = 3
        x_full = torch.from_numpy(np.random.default_rng(7).standard_normal((n, d)).astype(np.float32))
        x_owned = x_full[part.lo:part.hi].clone().requires_grad_(True)
        x_local = D._HaloFunction.apply(x_owned, ex)
        assert torch.equal(x_local.detach(), x_full[torch.from_numpy(part.local_to_global())])
        # the real product primitive refuses CPU tensors (no fallback)
        try:
            D.HaloExchange(part, "cpu").forward(x_owned.detach())
            raise AssertionError("CPU halo exchange must raise")
        except ngpde.NgpdeError:
            pass
        # transpose property: <halo(x), w> summed over ranks == <x, halo^T(w)> summed over ranks, and the pullback equals
        # the gradient of the single-process gather
        w_full = torch.from_numpy(np.random.default_rng(11 + rank).standard_normal((part.n_local, d)).astype(np.float32))
        (x_local * w_full).sum().backward()
        # single-process reference: every rank's local rows are a gather of x_full
        ws = [torch.zeros(1)] * world
        gathered = [None] * world
        dist.all_gather_object(gathered, (part.local_to_global(), w_full.numpy()))
        ref = torch.zeros(n, d, dtype=torch.float64)
        for l2g, w in gathered:
            ref.index_add_(0, torch.from_numpy(l2g), torch.from_numpy(w).double())
        got = x_owned.grad.double()
        assert torch.allclose(got, ref[part.lo:part.hi], rtol=1e-6, atol=1e-6), (got - ref[part.lo:part.hi]).abs().max()
        # gradient all-reduce helper
        g1, g2 = torch.full((5,), float(rank + 1)), torch.full((2, 2), 10.0 * (rank + 1))
        D.allreduce_gradients([g1, None, g2])
        tot = sum(range(1, world + 1))
        assert torch.equal(g1, torch.full((5,), float(tot))) and torch.equal(g2, torch.full((2, 2), 10.0 * tot))
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_halo_exchange_two_ranks_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


# ---- the C ABI's partitioner (csrc/ngpde_dist.cu) against the numpy statement: every array, element by element ----

@pytest.mark.parametrize("world", [1, 2, 5, 8])
@pytest.mark.parametrize("by", ["edges", "nodes"])
def test_native_partition_plan_matches_numpy(world, by):
    rng = np.random.default_rng(11 + world)
    cases = [(40, 300), (1, 0), (7, 1), (500, 6000), (64, 64)]
    for n, e in cases:
        s, t = rng.integers(0, n, e), rng.integers(0, n, e)
        for r in range(world):
            a = P.partition_nodes(s, t, n, world, r, by=by)
            b = P.partition_nodes_native(s, t, n, world, r, by=by)
            for f in a.__dataclass_fields__:
                va, vb = getattr(a, f), getattr(b, f)
                same = np.array_equal(va, vb) if isinstance(va, np.ndarray) else va == vb
                assert same, (n, e, world, r, by, f)
    # explicit bounds, including an empty rank
    n, e = 30, 200
    s, t = rng.integers(0, n, e), rng.integers(0, n, e)
    bounds = np.linspace(0, n, world + 1).astype(np.int64)
    if world > 2:
        bounds[2] = bounds[1]
    for r in range(world):
        a = P.partition_nodes(s, t, n, world, r, bounds=bounds)
        b = P.partition_nodes_native(s, t, n, world, r, bounds=bounds)
        assert all(np.array_equal(getattr(a, f), getattr(b, f)) for f in a.__dataclass_fields__)


def test_native_partition_rejects_bad_input():
    s, t = np.array([0, 5]), np.array([1, 1])
    with pytest.raises(ValueError):
        P.partition_nodes_native(s, t, 3, 2, 0)  # source 5 out of range
    with pytest.raises(ValueError):
        P.partition_nodes_native(np.array([0]), np.array([1]), 3, 2, 0, bounds=np.array([0, 4, 3]))


def test_morton_order_and_halo_locality():
    """An unsorted geometric graph: contiguous id ranges are random node sets (halo ~ everything); after the Morton
    renumbering they are compact blocks (halo ~ perimeter).  The C implementation equals the numpy statement."""
    rng = np.random.default_rng(5)
    n = 20000
    s, t, pos = workloads.radius_edges(n, 12.0, rng)
    scr = rng.permutation(n)
    s, t, _ = P.relabel(s, t, scr)
    pos = pos[:, scr]
    order = P.morton_order(pos)
    assert np.array_equal(np.sort(order), np.arange(n))
    assert np.array_equal(order, P.morton_order_numpy(pos))
    for dim in (1, 3):
        q = rng.uniform(-3, 5, size=(dim, 1000)).astype(np.float32)
        assert np.array_equal(P.morton_order(q), P.morton_order_numpy(q))
    halo_before = max(P.partition_nodes_native(s, t, n, 4, r).n_halo for r in range(4))
    s2, t2, inv = P.relabel(s, t, order)
    assert np.array_equal(order[s2], s) and np.array_equal(inv[order], np.arange(n))
    halo_after = max(P.partition_nodes_native(s2, t2, n, 4, r).n_halo for r in range(4))
    assert halo_before > 0.5 * n * 3 / 4 * 0.9 and halo_after < 0.06 * n, (halo_before, halo_after)


def test_partitioned_layer_refuses_gcnconv():
    from ngpde import distributed as D
    g = ngpde.GNNGraph([0, 1], [1, 0], num_nodes=2)
    with pytest.raises(TypeError):
        D.PartitionedLayer(ngpde.GCNConv((2, 2), initialgraph=g), g, 0, 1, "cpu")
