"""Multi-GPU execution of the message-passing path: one process per GPU over `torch.distributed` (NCCL over NVLink).

Two ways the path shards (index lists come from partition.py):

  * `PartitionedLayer` -- one large graph, nodes partitioned into contiguous ranges, owner-computes by destination.  Per
    layer call the boundary rows of `x` are exchanged (`HaloExchange`), the rank-local layer runs on [owned; halo] rows,
    and the pullback sends the halo cotangents home where they are added in a fixed order.  Static node data of the halo
    (positions, coefficients) is exchanged once at build time.  The forward result of every owned row is bit-identical
    to the single-GPU result (same messages, same order); parameter gradients are summed with one all-reduce.
  * `shard_ensemble` -- a block-diagonal batch of graphs: whole graphs per rank, no data-path collective; only the flat
    parameter gradient is all-reduced (`allreduce_gradients`).

Transfers: `mode="nccl"` packs with `ngpde_rows_gather` and calls `all_to_all_single`; `mode="put"` writes the rows
straight into the peers' halo buffers from inside the pack kernel (`ngpde_rows_put`, peer-mapped symmetric memory over
NVLink/NVSwitch), followed by a device-side barrier -- no send buffer, no NCCL call on the per-RHS path; `mode="native"`
runs the whole exchange inside the C ABI (`ngpde_halo_forward/backward`: pack kernel + grouped ncclSend/ncclRecv on a
communicator the library created from a unique id, csrc/ngpde_dist.cu) -- the path a Julia binder gets.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, ops
from .graph import GNNGraph, from_rowmajor, rowmajor
from .lux import NT
from .partition import (BatchShard, NodePartition, morton_order, partition_nodes, partition_nodes_native, relabel,
                        shard_batch)

Tensor = torch.Tensor


def _i32(a: np.ndarray, device) -> Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(device)


class NativeComm:
    """An NCCL communicator owned by libngpde (`ngpde_comm_init`): rank 0 draws the unique id through the C ABI, the id
    travels over the existing torch.distributed group (any side channel would do), every rank initialises on its current
    device.  From here on the data path needs no torch collective."""

    def __init__(self, device, group=None):
        self.lib = _lib.load()
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = torch.device(device)
        idbuf = (C.c_ubyte * _lib.UNIQUE_ID_BYTES)()
        if self.rank == 0:
            _lib.check(self.lib.ngpde_comm_unique_id(idbuf, None))
        backend = dist.get_backend(group)
        t = torch.tensor(list(idbuf), dtype=torch.uint8, device=self.device if backend == "nccl" else "cpu")
        src = dist.get_global_rank(group, 0) if group is not None else 0
        dist.broadcast(t, src=src, group=group)
        raw = bytes(t.cpu().tolist())
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ngpde_comm_init(C.byref(h), raw, self.world, self.rank, None))
        self.handle = h

    def allreduce_sum(self, t: Tensor) -> Tensor:
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
        with torch.cuda.device(t.device):
            _lib.check(self.lib.ngpde_allreduce_sum(self.handle, t.data_ptr(), t.numel(), ops._stream(t.device)))
        return t

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h is not None and h.value:
            try:
                self.lib.ngpde_comm_destroy(h)
            except Exception:
                pass


class PeerAllReduce:
    """Sum of a flat float32 vector over the ranks of one node without NCCL: the vector lives in a peer-mapped symmetric
    allocation, and after a device-side barrier every rank adds all `world` copies in ascending rank order with ONE kernel
    (`ngpde_peer_allreduce_sum`, loads over NVLink / NVSwitch).  Latency-bound sizes (the parameter gradient is ~100 KB)
    finish in a few microseconds, and the result is bit-identical on every rank.

        ar = PeerAllReduce(n, device)        # collective: every rank of the group
        runner = RhsRunner(..., dparams=ar.buffer)     # the backward writes the gradient straight into the symmetric buffer
        total = ar.reduce()                  # -> ar.out (local tensor of n floats)
    """

    def __init__(self, n: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.device = torch.device(device)
        grp = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(grp)
        self.n = int(n)
        npad = (self.n + 3) // 4 * 4
        self.buffer_full = symm_mem.empty(npad, dtype=torch.float32, device=self.device)
        self.buffer_full.zero_()
        self.hdl = symm_mem.rendezvous(self.buffer_full, grp)
        self.buffer = self.buffer_full[:self.n]
        self.out = torch.zeros(npad, dtype=torch.float32, device=self.device)[:self.n]
        self.peers = torch.tensor([int(p) for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=self.device)
        torch.cuda.synchronize(self.device)

    def reduce(self) -> Tensor:
        self.hdl.barrier(channel=0)  # every rank has written its buffer
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().ngpde_peer_allreduce_sum(self.peers.data_ptr(), self.world, self.out.data_ptr(), self.n,
                                                            ops._stream(self.device)))
        ops.LAUNCHES["count"] += 1
        self.hdl.barrier(channel=1)  # every rank has read: the buffers may be overwritten
        return self.out


class HaloExchange:
    """Per-RHS boundary exchange of one NodePartition.  `forward(x_owned [n_owned, d]) -> x_local [n_local, d]` and
    `backward(dx_local) -> dx_owned` are each other's transposes."""

    def __init__(self, part: NodePartition, device, group=None, mode: str = "nccl"):
        if mode not in ("nccl", "put", "native"):
            raise ValueError("mode must be 'nccl', 'put' or 'native'")
        self.part, self.device, self.group, self.mode = part, torch.device(device), group, mode
        self.send_rows = _i32(part.send_local, device)
        self.seg_rows, self.seg_ptr, self.seg_pos = (_i32(part.seg_rows, device), _i32(part.seg_ptr, device),
                                                     _i32(part.seg_pos, device))
        self.send_splits = [int(c) for c in part.send_counts]
        self.recv_splits = [int(c) for c in part.recv_counts]
        self.n_send = int(part.send_counts.sum())
        self._symm: Dict[int, tuple] = {}
        self._native = None

    # ---- native mode: plan + communicator + exchange all inside the C ABI ----
    def attach_native_plan(self, s: np.ndarray, t: np.ndarray, num_nodes: int, by: str = "edges"):
        """Build the native plan + communicator + halo object (mode='native'); called by PartitionedLayer."""
        lib, p = _lib.load(), self.part
        s = np.ascontiguousarray(s, dtype=np.int64)
        t = np.ascontiguousarray(t, dtype=np.int64)
        b = np.ascontiguousarray(p.bounds, dtype=np.int64)
        plan = C.c_void_p()
        _lib.check(lib.ngpde_partition_create(C.byref(plan), int(num_nodes), int(s.size), s.ctypes.data, t.ctypes.data,
                                              _lib.IDX_I64, 0, p.world, p.rank, 1 if by == "edges" else 0, b.ctypes.data))
        try:
            self.comm = NativeComm(self.device, self.group) if p.world > 1 else None
            h = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(lib.ngpde_halo_create(C.byref(h), plan, None if self.comm is None else self.comm.handle,
                                                 ops._stream(self.device)))
            self._native = h
        finally:
            lib.ngpde_partition_destroy(plan)

    def __del__(self):
        h, self._native = getattr(self, "_native", None), None
        if h is not None and h.value:
            try:
                _lib.load().ngpde_halo_destroy(h)
            except Exception:
                pass

    # ---- device primitives (CUDA only: the product path has no CPU fallback) ----
    def _pack(self, x: Tensor, rows: Tensor) -> Tensor:
        if not x.is_cuda:
            raise _lib.NgpdeError("halo packing runs on CUDA only (no CPU fallback)")
        out = torch.empty((rows.numel(), x.shape[1]), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().ngpde_rows_gather(x.data_ptr(), rows.data_ptr(), rows.numel(), x.shape[1],
                                                      out.data_ptr(), ops._stream(x.device)))
        ops.LAUNCHES["count"] += 1
        return out

    def _segment_add(self, dst: Tensor, src: Tensor) -> None:
        if not dst.is_cuda:
            raise _lib.NgpdeError("halo cotangent accumulation runs on CUDA only (no CPU fallback)")
        with torch.cuda.device(dst.device):
            _lib.check(_lib.load().ngpde_rows_segment_add(dst.data_ptr(), src.data_ptr(), self.seg_rows.data_ptr(),
                                                           self.seg_ptr.data_ptr(), self.seg_pos.data_ptr(),
                                                           self.seg_rows.numel(), dst.shape[1], ops._stream(dst.device)))
        ops.LAUNCHES["count"] += 1

    def _all_to_all(self, out: Tensor, inp: Tensor, out_splits: List[int], in_splits: List[int]) -> None:
        dist.all_to_all_single(out, inp, out_splits, in_splits, group=self.group)

    # ---- put mode: peer-mapped halo buffers ----
    def _symm_buffers(self, d: int):
        hit = self._symm.get(d)
        if hit is not None:
            return hit
        import torch.distributed._symmetric_memory as symm_mem
        p = self.part
        group = self.group if self.group is not None else dist.group.WORLD
        # every rank allocates the same size (symmetric): the largest halo of any rank
        nh = torch.tensor([p.n_halo], dtype=torch.int64, device=self.device)
        dist.all_reduce(nh, op=dist.ReduceOp.MAX, group=group)
        buf = symm_mem.empty((max(int(nh.item()), 1), d), dtype=torch.float32, device=self.device)
        hdl = symm_mem.rendezvous(buf, group)
        # where my rows start inside peer q's halo buffer = number of q's halo rows owned by ranks below me
        offs = torch.tensor([int(o) for o in p.peer_recv_offset], dtype=torch.int64, device=self.device)
        ptrs = [int(hdl.buffer_ptrs[q]) + 4 * d * int(p.peer_recv_offset[q]) for q in range(p.world)]
        peer_dst = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        peer_ptr = torch.tensor(np.concatenate([[0], np.cumsum(p.send_counts)]), dtype=torch.int64, device=self.device)
        self._symm[d] = (buf, hdl, peer_dst, peer_ptr, offs)
        return self._symm[d]

    def forward(self, x_owned: Tensor) -> Tensor:
        p = self.part
        d = x_owned.shape[1]
        x_local = torch.empty((p.n_local, d), dtype=torch.float32, device=x_owned.device)
        if self.mode == "native":
            if self._native is None:
                raise _lib.NgpdeError("mode='native': no native plan attached (build the exchange through PartitionedLayer)")
            with torch.cuda.device(x_owned.device):
                _lib.check(_lib.load().ngpde_halo_forward(self._native, x_owned.data_ptr(), d, x_local.data_ptr(),
                                                          ops._stream(x_owned.device)))
            ops.LAUNCHES["count"] += 1
            return x_local
        x_local[:p.n_owned].copy_(x_owned)
        if p.world == 1:
            return x_local
        if self.mode == "put":
            buf, hdl, peer_dst, peer_ptr, _ = self._symm_buffers(d)
            hdl.barrier(channel=0)  # every peer has finished reading the previous halo
            with torch.cuda.device(x_owned.device):
                _lib.check(_lib.load().ngpde_rows_put(x_owned.data_ptr(), self.send_rows.data_ptr(), peer_ptr.data_ptr(),
                                                       peer_dst.data_ptr(), p.world, self.n_send, d,
                                                       ops._stream(x_owned.device)))
            ops.LAUNCHES["count"] += 1
            hdl.barrier(channel=1)  # all rows addressed to me have landed
            x_local[p.n_owned:].copy_(buf[:p.n_halo])
        else:
            send = self._pack(x_owned, self.send_rows)
            self._all_to_all(x_local[p.n_owned:], send, self.recv_splits, self.send_splits)
        return x_local

    def backward(self, dx_local: Tensor) -> Tensor:
        p = self.part
        if self.mode == "native":
            dx_owned = torch.empty((p.n_owned, dx_local.shape[1]), dtype=torch.float32, device=dx_local.device)
            with torch.cuda.device(dx_local.device):
                _lib.check(_lib.load().ngpde_halo_backward(self._native, dx_local.data_ptr(), dx_local.shape[1],
                                                           dx_owned.data_ptr(), ops._stream(dx_local.device)))
            ops.LAUNCHES["count"] += 1
            return dx_owned
        dx_owned = dx_local[:p.n_owned].clone()
        if p.world == 1:
            return dx_owned
        back = torch.empty((self.n_send, dx_local.shape[1]), dtype=torch.float32, device=dx_local.device)
        self._all_to_all(back, dx_local[p.n_owned:].contiguous(), self.send_splits, self.recv_splits)
        self._segment_add(dx_owned, back)
        return dx_owned


class _HaloFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_owned: Tensor, ex: HaloExchange):
        ctx.ex = ex
        return ex.forward(x_owned.contiguous())

    @staticmethod
    def backward(ctx, g: Tensor):
        return ctx.ex.backward(g.contiguous()), None


def exchange_static(ex: HaloExchange, a_owned: Tensor) -> Tensor:
    """One-time exchange of static node data (graph `ndata`): [n_owned, d] -> [n_local, d], no gradient."""
    saved, ex.mode = ex.mode, ("native" if ex.mode == "native" else "nccl")
    try:
        with torch.no_grad():
            return ex.forward(a_owned.contiguous())
    finally:
        ex.mode = saved


class PartitionedLayer:
    """A message-passing layer over one node-partitioned graph.

        pl = PartitionedLayer(layer, g_full, rank, world, device)    # g_full: the whole GNNGraph (host or device)
        y_owned, st = pl(x_owned, ps, st_local)                      # x_owned: (d, n_owned) Julia-shaped

    `pl.st` is the rank-local state (graph = local topology + [owned; halo] node data + local edge data).
    """

    def __init__(self, layer, g_full: GNNGraph, rank: int, world: int, device, group=None, mode: str = "nccl",
                 by: str = "edges", exchange_cls=HaloExchange, order=None):
        if not hasattr(layer, "prepare"):
            # GCNConv (alone or inside a Chain) normalises by the in-degree of the SOURCE as well: halo rows have no local
            # in-edges, so their 1/sqrt(deg) factor would be wrong and the owned outputs silently incorrect
            raise TypeError("PartitionedLayer shards ExplicitEdgeConv / VMHConv / MPPDEConv / GNOConv; GCNConv needs the halo's "
                            "global in-degrees and is sharded by whole graphs instead (shard_ensemble)")
        self.layer, self.rank, self.world = layer, rank, world
        self.device = torch.device(device)
        s, t = g_full.s.cpu().numpy(), g_full.t.cpu().numpy()
        # optional spatial renumbering: contiguous id ranges become compact blocks of the Morton curve, so a graph whose
        # node numbering has no locality still gets an O(sqrt(N)) halo
        self.order: Optional[np.ndarray] = None
        if isinstance(order, str):
            if order != "morton":
                raise ValueError("order must be 'morton', a permutation, or None")
            if "x" not in g_full.ndata:
                raise KeyError("order='morton' needs the node coordinates in g.ndata.x")
            order = morton_order(g_full.ndata["x"].cpu().numpy())
        if order is not None:
            self.order = np.asarray(order, dtype=np.int64)
            s, t, _ = relabel(s, t, self.order)
        self.part = partition_nodes_native(s, t, g_full.num_nodes, world, rank, by=by)
        p = self.part
        self.exchange = exchange_cls(p, self.device, group, mode)
        if mode == "native":
            self.exchange.attach_native_plan(s, t, g_full.num_nodes, by)
        l2g = p.local_to_global()
        if self.order is not None:
            l2g = self.order[l2g]
        l2g = torch.from_numpy(l2g)
        eid = torch.from_numpy(p.edge_ids)
        # rank-local copies of the static data: owned + halo node rows (a slice of the host copy every rank built from the
        # same seed; `exchange_static` does the same over the wire when only owned rows are at hand), local edge rows
        nd = {k: v.cpu()[:, l2g].to(self.device) for k, v in g_full.ndata.items()}
        ed = {k: v.cpu()[:, eid].to(self.device) for k, v in g_full.edata.items()}
        self.graph = GNNGraph(torch.from_numpy(p.s_local), torch.from_numpy(p.t_local), num_nodes=p.n_local, ndata=nd,
                              edata=ed, gdata=g_full.gdata).to(self.device)
        self.owned_global = l2g[:p.n_owned]  # ids (in the caller's numbering) of the rows this rank owns, in local order

    def local_state(self, st: NT) -> NT:
        from .utils import updategraph
        return updategraph(st, self.graph)

    def owned(self, x_full: Tensor) -> Tensor:
        """Columns of a full (d, N) array owned by this rank (in local row order)."""
        if self.order is None:
            return x_full[:, self.part.lo:self.part.hi]
        return x_full[:, self.owned_global.to(x_full.device)]

    def __call__(self, x_owned: Tensor, ps, st_local: NT):
        x_rm = rowmajor(x_owned)
        x_local = _HaloFunction.apply(x_rm, self.exchange)
        y_local, st_out = self.layer(from_rowmajor(x_local), ps, st_local)
        return y_local[:, :self.part.n_owned], st_out


def shard_ensemble(g_full: GNNGraph, rank: int, world: int, device) -> tuple:
    """Rank-local sub-batch of a block-diagonal batch (whole graphs [g0, g1)); returns (GNNGraph, BatchShard)."""
    s, t = g_full.s.cpu().numpy(), g_full.t.cpu().numpy()
    sh = shard_batch(s, t, g_full.num_nodes, g_full.num_graphs, world, rank)
    eid = torch.from_numpy(sh.edge_ids)
    nd = {k: v.cpu()[:, sh.node_lo:sh.node_hi] for k, v in g_full.ndata.items()}
    ed = {k: v.cpu()[:, eid] for k, v in g_full.edata.items()}
    gd = {k: (v if v.dim() == 2 else v.reshape(-1, 1)).cpu()[:, sh.g0:sh.g1] for k, v in g_full.gdata.items()}
    g = GNNGraph(torch.from_numpy(sh.s_local), torch.from_numpy(sh.t_local), num_nodes=sh.node_hi - sh.node_lo, ndata=nd,
                 edata=ed, gdata=gd, num_graphs=sh.g1 - sh.g0).to(device)
    return g, sh


def allreduce_gradients(grads: Sequence[Optional[Tensor]], group=None) -> None:
    """Sum the flat parameter gradients over ranks, one collective for all of them (they are a few hundred KB at most:
    latency-bound, so they are coalesced into one buffer)."""
    gs = [g for g in grads if g is not None]
    if not gs or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    if len(gs) == 1 and gs[0].is_contiguous():
        dist.all_reduce(gs[0], group=group)
        return
    flat = torch.cat([g.reshape(-1) for g in gs])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in gs:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
