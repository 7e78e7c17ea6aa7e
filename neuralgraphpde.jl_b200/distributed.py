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
NVLink/NVSwitch), followed by a device-side barrier -- no send buffer, no NCCL call on the per-RHS path.
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
from .partition import BatchShard, NodePartition, partition_nodes, shard_batch

Tensor = torch.Tensor


def _i32(a: np.ndarray, device) -> Tensor:
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(device)


class HaloExchange:
    """Per-RHS boundary exchange of one NodePartition.  `forward(x_owned [n_owned, d]) -> x_local [n_local, d]` and
    `backward(dx_local) -> dx_owned` are each other's transposes."""

    def __init__(self, part: NodePartition, device, group=None, mode: str = "nccl"):
        if mode not in ("nccl", "put"):
            raise ValueError("mode must be 'nccl' or 'put'")
        self.part, self.device, self.group, self.mode = part, torch.device(device), group, mode
        self.send_rows = _i32(part.send_local, device)
        self.seg_rows, self.seg_ptr, self.seg_pos = (_i32(part.seg_rows, device), _i32(part.seg_ptr, device),
                                                     _i32(part.seg_pos, device))
        self.send_splits = [int(c) for c in part.send_counts]
        self.recv_splits = [int(c) for c in part.recv_counts]
        self.n_send = int(part.send_counts.sum())
        self._symm: Dict[int, tuple] = {}

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
    saved, ex.mode = ex.mode, "nccl"
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
                 by: str = "edges", exchange_cls=HaloExchange):
        self.layer, self.rank, self.world = layer, rank, world
        self.device = torch.device(device)
        s, t = g_full.s.cpu().numpy(), g_full.t.cpu().numpy()
        self.part = partition_nodes(s, t, g_full.num_nodes, world, rank, by=by)
        p = self.part
        self.exchange = exchange_cls(p, self.device, group, mode)
        l2g = torch.from_numpy(p.local_to_global())
        eid = torch.from_numpy(p.edge_ids)
        # rank-local copies of the static data: owned + halo node rows (a slice of the host copy every rank built from the
        # same seed; `exchange_static` does the same over the wire when only owned rows are at hand), local edge rows
        nd = {k: v.cpu()[:, l2g].to(self.device) for k, v in g_full.ndata.items()}
        ed = {k: v.cpu()[:, eid].to(self.device) for k, v in g_full.edata.items()}
        self.graph = GNNGraph(torch.from_numpy(p.s_local), torch.from_numpy(p.t_local), num_nodes=p.n_local, ndata=nd,
                              edata=ed, gdata=g_full.gdata).to(self.device)

    def local_state(self, st: NT) -> NT:
        from .utils import updategraph
        return updategraph(st, self.graph)

    def owned(self, x_full: Tensor) -> Tensor:
        """Columns of a full (d, N) array owned by this rank."""
        return x_full[:, self.part.lo:self.part.hi]

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
    flat = torch.cat([g.reshape(-1) for g in gs])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in gs:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
