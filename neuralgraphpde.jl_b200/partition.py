"""Host side of the multi-GPU message-passing path: how one graph, or one batch of graphs, is
split over ranks.  Pure numpy, no device work; the index lists produced here are part of the bit-exact index contract
and are what `distributed.HaloExchange` feeds to the pack / segment-add kernels and to the collective.

The reference has no distributed code (no NCCL/MPI call site under /root/reference); what it fixes is the semantics a
partition must preserve: `propagate` gathers `x[s]`, `x[t]` per edge and reduces at `t` in stored edge order
(/root/reference/src/layers.jl:111,326,416,534 through GraphNeuralNetworks [DEP]).  Hence

  * owner-computes by destination: rank r owns the contiguous node range [bounds[r], bounds[r+1]) and stores every edge
    whose target it owns, in the original relative order -- so each owned row reduces exactly the same messages in
    exactly the same order as on one GPU and the forward result is bit-identical;
  * sources owned elsewhere form the halo, appended after the owned rows in ascending global id (= grouped by owner);
  * batched ensembles (block-diagonal graphs, MPPDEConv batches layers.jl:361,394) shard by whole graphs: no halo.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np


def balanced_bounds(t: np.ndarray, num_nodes: int, world: int, by: str = "edges") -> np.ndarray:
    """Contiguous node ranges per rank.  by="nodes": equal node counts; by="edges": equal in-edge counts (the work of the
    edge phase), cut positions rounded to node boundaries by a lower-bound search on the in-degree prefix sum."""
    if world < 1:
        raise ValueError("world must be >= 1")
    if by == "nodes" or len(t) == 0:
        return (np.arange(world + 1, dtype=np.int64) * num_nodes) // world
    if by != "edges":
        raise ValueError(f"unknown balancing criterion {by!r}")
    deg = np.bincount(np.asarray(t, dtype=np.int64), minlength=num_nodes).astype(np.int64)
    # weight a node by its in-edges plus one (the node phase), so that edgeless stretches are still spread out
    cum = np.concatenate([[0], np.cumsum(deg + 1)])
    targets = (np.arange(1, world, dtype=np.int64) * cum[-1]) // world
    cuts = np.searchsorted(cum, targets, side="left")
    b = np.concatenate([[0], cuts, [num_nodes]]).astype(np.int64)
    return np.maximum.accumulate(b)


@dataclass
class NodePartition:
    """Rank-local view of a node-partitioned graph.  Local node ids: [0, n_owned) owned (global id - lo), then the halo."""
    rank: int
    world: int
    bounds: np.ndarray          # [world+1] global node ranges
    lo: int
    hi: int
    halo_global: np.ndarray     # [n_halo] global ids of imported source nodes, ascending (grouped by owner)
    recv_counts: np.ndarray     # [world] halo rows received from each peer (recv_counts[rank] == 0)
    send_counts: np.ndarray     # [world] owned rows sent to each peer
    send_local: np.ndarray      # [sum send_counts] owned-local row ids, grouped by peer, each group ascending
    s_local: np.ndarray         # [E_local] local source ids (owned or halo)
    t_local: np.ndarray         # [E_local] local target ids (always owned)
    edge_ids: np.ndarray        # [E_local] original COO positions, ascending
    # backward: halo cotangents coming home are added per owned row over ascending position in the receive buffer
    seg_rows: np.ndarray        # [U] distinct owned-local rows that have at least one remote reader
    seg_ptr: np.ndarray         # [U+1]
    seg_pos: np.ndarray         # [sum send_counts] positions in the (peer-major) receive buffer
    peer_recv_offset: np.ndarray  # [world] row offset of my rows inside peer p's halo block (for direct peer stores)

    @property
    def n_owned(self) -> int:
        return self.hi - self.lo

    @property
    def n_halo(self) -> int:
        return int(self.halo_global.size)

    @property
    def n_local(self) -> int:
        return self.n_owned + self.n_halo

    def local_to_global(self) -> np.ndarray:
        return np.concatenate([np.arange(self.lo, self.hi, dtype=np.int64), self.halo_global])


def _halo_of(s: np.ndarray, t: np.ndarray, lo: int, hi: int):
    mine = (t >= lo) & (t < hi)
    eid = np.nonzero(mine)[0]
    src = s[eid]
    remote = src[(src < lo) | (src >= hi)]
    return eid, np.unique(remote)


def partition_nodes(s: np.ndarray, t: np.ndarray, num_nodes: int, world: int, rank: int,
                    bounds: Optional[np.ndarray] = None, by: str = "edges") -> NodePartition:
    """Partition for `rank`, computed from the full edge list (every rank holds it at build time; nothing here is
    communicated).  0-based indices."""
    s = np.asarray(s, dtype=np.int64)
    t = np.asarray(t, dtype=np.int64)
    if bounds is None:
        bounds = balanced_bounds(t, num_nodes, world, by)
    bounds = np.asarray(bounds, dtype=np.int64)
    if bounds.shape != (world + 1,) or bounds[0] != 0 or bounds[-1] != num_nodes or np.any(np.diff(bounds) < 0):
        raise ValueError("bounds must be a non-decreasing [world+1] vector from 0 to num_nodes")
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    eid, halo = _halo_of(s, t, lo, hi)
    owner = np.searchsorted(bounds, halo, side="right") - 1
    recv_counts = np.bincount(owner, minlength=world).astype(np.int64)
    # local ids of the sources: owned -> id - lo; halo -> n_owned + rank within the sorted halo list
    src = s[eid]
    is_own = (src >= lo) & (src < hi)
    s_local = np.where(is_own, src - lo, (hi - lo) + np.searchsorted(halo, src))
    t_local = t[eid] - lo
    # what every peer imports from me: the same computation from the peer's point of view
    send_lists: List[np.ndarray] = []
    peer_off = np.zeros(world, dtype=np.int64)
    for p in range(world):
        if p == rank or hi == lo:
            send_lists.append(np.zeros(0, dtype=np.int64))
            continue
        _, halo_p = _halo_of(s, t, int(bounds[p]), int(bounds[p + 1]))
        send_lists.append(halo_p[(halo_p >= lo) & (halo_p < hi)] - lo)
        peer_off[p] = np.searchsorted(halo_p, lo, side="left")  # p's halo rows owned by ranks below me come first
    send_counts = np.array([len(x) for x in send_lists], dtype=np.int64)
    send_local = np.concatenate(send_lists) if send_lists else np.zeros(0, dtype=np.int64)
    order = np.argsort(send_local, kind="stable")
    rows_sorted = send_local[order]
    if rows_sorted.size:
        head = np.concatenate([[True], rows_sorted[1:] != rows_sorted[:-1]])
        seg_rows = rows_sorted[head]
        seg_ptr = np.concatenate([np.nonzero(head)[0], [rows_sorted.size]])
    else:
        seg_rows = np.zeros(0, dtype=np.int64)
        seg_ptr = np.zeros(1, dtype=np.int64)
    return NodePartition(rank, world, bounds, lo, hi, halo, recv_counts, send_counts, send_local, s_local, t_local, eid,
                         seg_rows, seg_ptr.astype(np.int64), order.astype(np.int64), peer_off)


@dataclass
class BatchShard:
    """Whole graphs [g0, g1) of a block-diagonal batch of `num_graphs` equal-sized graphs."""
    rank: int
    world: int
    g0: int
    g1: int
    node_lo: int
    node_hi: int
    s_local: np.ndarray
    t_local: np.ndarray
    edge_ids: np.ndarray


def shard_batch(s: np.ndarray, t: np.ndarray, num_nodes: int, num_graphs: int, world: int, rank: int) -> BatchShard:
    """Contiguous ranges of whole graphs per rank (graphs are laid out contiguously and have equal size -- the same
    assumption MPPDEConv makes at layers.jl:410,418).  Edges keep their original relative order."""
    if num_graphs < 1 or num_nodes % num_graphs != 0:
        raise ValueError(f"a batch of {num_graphs} equal-sized graphs cannot hold {num_nodes} nodes")
    per = num_nodes // num_graphs
    g0, g1 = (rank * num_graphs) // world, ((rank + 1) * num_graphs) // world
    lo, hi = g0 * per, g1 * per
    s = np.asarray(s, dtype=np.int64)
    t = np.asarray(t, dtype=np.int64)
    mine = (t >= lo) & (t < hi)
    eid = np.nonzero(mine)[0]
    src = s[eid]
    if src.size and (src.min() < lo or src.max() >= hi):
        raise ValueError("the batch is not block-diagonal: an edge crosses graphs owned by different ranks")
    return BatchShard(rank, world, g0, g1, lo, hi, src - lo, t[eid] - lo, eid)


def partition_nodes_native(s: np.ndarray, t: np.ndarray, num_nodes: int, world: int, rank: int,
                           bounds: Optional[np.ndarray] = None, by: str = "edges") -> NodePartition:
    """The same plan built by the C ABI (`ngpde_partition_create`, csrc/ngpde_dist.cu) -- what `PartitionedLayer` uses and
    what a Julia binder calls; `partition_nodes` above is its numpy statement, kept as the checker of the CPU tests."""
    import ctypes as C
    from . import _lib
    if by not in ("edges", "nodes"):
        raise ValueError(f"unknown balancing criterion {by!r}")
    lib = _lib.load()
    s = np.ascontiguousarray(s, dtype=np.int64)
    t = np.ascontiguousarray(t, dtype=np.int64)
    b = None if bounds is None else np.ascontiguousarray(bounds, dtype=np.int64)
    if b is not None and b.shape != (world + 1,):
        raise ValueError("bounds must be a non-decreasing [world+1] vector from 0 to num_nodes")
    h = C.c_void_p()
    rc = lib.ngpde_partition_create(C.byref(h), int(num_nodes), int(s.size), s.ctypes.data, t.ctypes.data, _lib.IDX_I64, 0,
                                    int(world), int(rank), 1 if by == "edges" else 0, None if b is None else b.ctypes.data)
    if rc != 0:
        msg = lib.ngpde_last_error().decode("utf-8", "replace")
        raise ValueError(msg) if rc == -1 else _lib.NgpdeError(f"libngpde error {rc}: {msg}")
    try:
        def arr(name):
            ptr, n = C.c_void_p(), C.c_int64()
            _lib.check(lib.ngpde_partition_array(h, _lib.PA[name], C.byref(ptr), C.byref(n)))
            if n.value == 0:
                return np.zeros(0, dtype=np.int64)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int64)), shape=(n.value,)).copy()
        a = {k: arr(k) for k in _lib.PA}
    finally:
        lib.ngpde_partition_destroy(h)
    bd = a["bounds"]
    return NodePartition(rank, world, bd, int(bd[rank]), int(bd[rank + 1]), a["halo_global"], a["recv_counts"],
                         a["send_counts"], a["send_local"], a["s_local"], a["t_local"], a["edge_ids"], a["seg_rows"],
                         a["seg_ptr"], a["seg_pos"], a["peer_recv_offset"])


def morton_order(pos: np.ndarray) -> np.ndarray:
    """Z-order permutation of the nodes from coordinates `pos` (dim, N) Julia-shaped or (N, dim): order[k] = id of the k-th
    node along the curve (C ABI `ngpde_morton_order`)."""
    from . import _lib
    p = np.asarray(pos, dtype=np.float32)
    if p.ndim == 1:
        p = p[None, :]
    if p.shape[0] <= 3 and p.shape[1] > 3:
        p = p.T
    p = np.ascontiguousarray(p)
    order = np.empty(p.shape[0], dtype=np.int64)
    _lib.check(_lib.load().ngpde_morton_order(p.ctypes.data, p.shape[0], p.shape[1], order.ctypes.data))
    return order


def morton_order_numpy(pos: np.ndarray) -> np.ndarray:
    """numpy statement of `morton_order` (test checker)."""
    p = np.asarray(pos, dtype=np.float32)
    if p.ndim == 1:
        p = p[None, :]
    if p.shape[0] <= 3 and p.shape[1] > 3:
        p = p.T
    n, dim = p.shape
    mn, mx = p.min(axis=0).astype(np.float64), p.max(axis=0).astype(np.float64)
    span = mx - mn
    u = np.where(span > 0, (p.astype(np.float64) - mn) / np.where(span > 0, span, 1.0), 0.0)
    q = np.minimum(np.floor(u * float(1 << 21)), float((1 << 21) - 1)).astype(np.uint64)
    code = np.zeros(n, dtype=np.uint64)
    for bit in range(20, -1, -1):
        for a in range(dim):
            code = (code << np.uint64(1)) | ((q[:, a] >> np.uint64(bit)) & np.uint64(1))
    return np.argsort(code, kind="stable").astype(np.int64)


def relabel(s: np.ndarray, t: np.ndarray, order: np.ndarray):
    """Edge lists under the renumbering new_id = rank of the node in `order` (edges keep their stored order, so every
    destination still reduces the same messages in the same order).  Returns (s_new, t_new, inverse) with
    inverse[old_id] = new_id; node-indexed arrays move as `a[..., order]`."""
    inv = np.empty(order.size, dtype=np.int64)
    inv[order] = np.arange(order.size, dtype=np.int64)
    return inv[np.asarray(s, dtype=np.int64)], inv[np.asarray(t, dtype=np.int64)], inv


def partition_summary(parts: List[NodePartition]) -> Dict[str, float]:
    """Balance and halo statistics over all ranks (for logs and DESIGN.md tables)."""
    e = np.array([p.edge_ids.size for p in parts], dtype=np.float64)
    n = np.array([p.n_owned for p in parts], dtype=np.float64)
    h = np.array([p.n_halo for p in parts], dtype=np.float64)
    return {"edges_max_over_mean": float(e.max() / max(e.mean(), 1.0)), "nodes_max_over_mean": float(n.max() / max(n.mean(), 1.0)),
            "halo_rows_max": float(h.max()), "halo_frac_max": float((h / np.maximum(n, 1.0)).max())}
