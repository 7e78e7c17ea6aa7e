"""Preallocated right-hand-side runner: one message-passing layer bound to a graph, its buffers and workspace, calling
the C ABI directly (no autograd bookkeeping, no allocator traffic) -- what an ODE integration loop or a benchmark calls
thousands of times.  `forward()` evaluates y = layer(x); `backward()` the VJP w.r.t. (x, ps) for the cotangent in
`self.dy`.  Optionally the forward+backward pair is captured once into a CUDA graph and replayed.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib, ops
from .lux import ComponentArray

Tensor = torch.Tensor


class RhsRunner:
    def __init__(self, layer, x: Tensor, ps, st, use_cuda_graph: bool = False, share: Optional["RhsRunner"] = None,
                 dparams: Optional[Tensor] = None, use_state: bool = True):
        """`share`: another runner of the same layer / graph whose parameter buffers and workspaces this one uses too
        (the stages of one Runge-Kutta step: same parameters, calls strictly one after the other)."""
        if not hasattr(layer, "prepare"):
            raise TypeError("RhsRunner binds one of ExplicitEdgeConv / VMHConv / MPPDEConv / GNOConv")
        self.lib = _lib.load()
        self._layer, self._st = layer, st
        (x_rm, phi, node, self.handle, self.desc, self.snode, self.edata, self.theta, dm, dy) = layer.prepare(x, ps, st)
        dev = x_rm.device
        self.dev = dev
        self.x = x_rm.detach().clone()
        if share is not None:
            self.phi, self.node = share.phi, share.node
        else:
            self.phi = phi.detach().contiguous().clone()
            self.node = None if node is None else node.detach().contiguous().clone()
        self.has_node = self.desc.node.n_layers > 0
        N = self.x.shape[0]
        self.mbar = torch.empty((N, dm), dtype=torch.float32, device=dev)
        self.y = torch.empty((N, dy), dtype=torch.float32, device=dev) if self.has_node else self.mbar
        self.dy = torch.zeros_like(self.y)
        self.dx = torch.empty_like(self.x)
        # one flat gradient buffer [dphi | dnode] (phi padded to a 16-byte boundary): a data-parallel step all-reduces it
        # with ONE collective, no staging copy
        nphi = self.phi.numel()
        pad = (-nphi) % 4
        n_dp = nphi + pad + (0 if self.node is None else self.node.numel())
        if dparams is not None:  # caller-provided gradient buffer (e.g. a peer-mapped symmetric allocation)
            if dparams.numel() != n_dp or dparams.dtype != torch.float32 or not dparams.is_contiguous():
                raise ValueError(f"dparams must be a contiguous float32 vector of {n_dp} entries")
            self.dparams = dparams
        else:
            self.dparams = torch.zeros(n_dp, dtype=torch.float32, device=dev)
        self.dphi = self.dparams[:nphi]
        self.dnode = None if self.node is None else self.dparams[nphi + pad:]
        with torch.cuda.device(dev):
            nbytes = self.lib.ngpde_conv_workspace_bytes(self.handle, C.byref(self.desc), 1)
            if nbytes == 0:
                _lib.check(-1)
            nbytes_f = self.lib.ngpde_conv_workspace_bytes(self.handle, C.byref(self.desc), 0)
            if nbytes_f == 0:
                _lib.check(-1)
        if share is not None and share.ws.numel() >= nbytes and share.ws_fwd.numel() >= nbytes_f:
            self.ws, self.ws_fwd = share.ws, share.ws_fwd
        else:
            self.ws = torch.empty(int(nbytes), dtype=torch.uint8, device=dev)
            self.ws_fwd = torch.empty(int(nbytes_f), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            nstate = self.lib.ngpde_conv_state_bytes(self.handle, C.byref(self.desc))
        # forward -> backward state (hoisted first-layer projections): private to this runner, like mbar
        # (use_state=False: io.state stays NULL and the backward recomputes what it needs -- same results, less memory)
        self.state = torch.empty(int(nstate), dtype=torch.uint8, device=dev) if (nstate and use_state) else None
        p = ops._ptr
        self.io = _lib.ConvIO(x=p(self.x), snode=p(self.snode), edata=p(self.edata), theta=p(self.theta),
                              phi_params=p(self.phi), node_params=p(self.node), mbar=p(self.mbar), y=p(self.y),
                              dy=p(self.dy), dx=p(self.dx), dphi_params=p(self.dphi), dnode_params=p(self.dnode),
                              state=p(self.state))
        fam = ops._FAMILY_FN[self.desc.family]
        self._fwd = getattr(self.lib, f"ngpde_{fam}_forward")
        self._bwd = getattr(self.lib, f"ngpde_{fam}_backward")
        factored = ops._factored(self.lib, self.handle, self.desc)  # GNOConv: + GEMMs, cotangent scale, split-K reduce
        self.launches_fwd = (2 if self.has_node else 1) + (1 if factored else 0)
        self.launches_bwd = (8 if self.has_node else 5) + (4 if factored else 0)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.graph_kernels: Optional[int] = None  # kernel nodes of the captured step (exact launch count per replay)
        # Everything a captured graph points at is owned by this object (x / parameters are private copies, the packed
        # static data and the native graph layout are referenced from here), so replaying stays valid whatever happens
        # to the caller's ps / st afterwards; new values come in through `set_params` / `x.copy_`.
        self._keep = (self.snode, self.edata, self.theta, self.handle)
        if use_cuda_graph:
            self.capture()

    @staticmethod
    def dparams_len(layer, x: Tensor, ps, st) -> int:
        """Length of the flat gradient buffer [dphi | pad | dnode] a runner of this layer uses."""
        pr = layer.prepare(x, ps, st)
        nphi = pr[1].numel()
        return nphi + ((-nphi) % 4) + (0 if pr[2] is None else pr[2].numel())

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.dev).cuda_stream

    def forward(self) -> Tensor:
        _lib.check(self._fwd(self.handle, C.byref(self.desc), C.byref(self.io), self.ws_fwd.data_ptr(), self.ws_fwd.numel(),
                             self._stream()))
        ops.LAUNCHES["count"] += self.launches_fwd
        return self.y

    def backward(self):
        _lib.check(self._bwd(self.handle, C.byref(self.desc), C.byref(self.io), self.ws.data_ptr(), self.ws.numel(),
                             self._stream()))
        ops.LAUNCHES["count"] += self.launches_bwd
        return self.dx, self.dphi, self.dnode

    def set_params(self, ps) -> None:
        """Copy new parameter values (same tree) into the bound buffers, e.g. after an optimiser step."""
        pr = self._layer.prepare(self.x.T, ps, self._st)
        self.phi.copy_(pr[1].detach())
        if self.node is not None:
            self.node.copy_(pr[2].detach())

    def capture(self, extra=None):
        """Capture forward+backward (and `extra()`, e.g. the gradient all-reduce) into one CUDA graph."""
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            self.forward()
            self.backward()
            if extra is not None:
                extra()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        l0 = ops.LAUNCHES["count"]
        self.graph = torch.cuda.CUDAGraph(keep_graph=True)
        with torch.cuda.graph(self.graph):
            self.forward()
            self.backward()
            if extra is not None:
                extra()
        ops.LAUNCHES["count"] = l0  # a capture launches nothing
        try:
            self.graph_kernels, _ = _lib.cuda_graph_kernel_nodes(self.graph.raw_cuda_graph())
        except Exception:  # noqa: BLE001 -- counting is a diagnostic; the table below stays as the estimate
            self.graph_kernels = None
        self.graph.instantiate()

    def step(self):
        """One RHS evaluation, forward + VJP."""
        if self.graph is not None:
            self.graph.replay()
            ops.LAUNCHES["count"] += (self.graph_kernels if self.graph_kernels is not None
                                      else self.launches_fwd + self.launches_bwd)
        else:
            self.forward()
            self.backward()


class ChainRhsRunner:
    """Forward + VJP of a Lux `Chain` of graph layers (C5: GCNConv -> GCNConv -> VMHConv) through the autograd bridge: every
    layer call and every pullback goes through the C ABI (`ops.ConvFunction` / `ops.GcnFunction`), torch only links them.
    Same surface as `RhsRunner` where `bench.py` needs it (`y`, `dy`, `step()`, `grads`, `handle`/`desc` of the last
    message-passing layer for the kernel-path query)."""

    def __init__(self, layer, x: Tensor, ps, st):
        self.layer, self.st = layer, st
        self.ca = ComponentArray(ps)
        self.ca.data.requires_grad_(True)
        self.x = x.detach().clone().requires_grad_(True)
        # one eager pass: output shape, and the descriptor of the last layer that has a fused MLP
        h = self.x.detach()
        self.handle = self.desc = None
        for i, l in enumerate(layer.layers):
            k = f"layer_{i + 1}"
            if hasattr(l, "prepare"):
                pr = l.prepare(h, getattr(ps, k), st[k])
                self.handle, self.desc = pr[3], pr[4]
            h, _ = l(h, getattr(ps, k), st[k])
        self.y = h.detach().T.contiguous()  # [N, d], the row-major image of Julia's (d, N)
        self.dy = torch.zeros_like(self.y)
        self.grads = [self.ca.data]

    def step(self):
        self.x.grad = None
        self.ca.data.grad = None
        y, _ = self.layer(self.x, self.ca, self.st)
        y.backward(self.dy.T)
        return y

    @property
    def dparams(self) -> Tensor:
        return self.ca.data.grad


class PartitionedRhsRunner:
    """RhsRunner over this rank's share of a node-partitioned graph (distributed.PartitionedLayer): every step exchanges
    the boundary rows of x, runs the local forward + VJP, sends the halo cotangents home and all-reduces the flat
    parameter gradients -- the per-RHS sequence of SURVEY.md section 8e."""

    def __init__(self, pl, x_owned: Tensor, ps, st):
        from .distributed import allreduce_gradients
        from .graph import from_rowmajor, rowmajor
        self.pl, self._allreduce = pl, allreduce_gradients
        p = pl.part
        self.n_owned = p.n_owned
        self.x_owned = rowmajor(x_owned).detach().clone()
        x_local = pl.exchange.forward(self.x_owned)
        self.runner = RhsRunner(pl.layer, from_rowmajor(x_local), ps, pl.local_state(st))
        self.dy_owned = torch.zeros((p.n_owned, self.runner.dy.shape[1]), dtype=torch.float32, device=self.x_owned.device)
        self.dx_owned: Optional[Tensor] = None

    @property
    def y_owned(self) -> Tensor:
        return self.runner.y[:self.n_owned]

    def step(self):
        r = self.runner
        r.x.copy_(self.pl.exchange.forward(self.x_owned))
        r.forward()
        r.dy[self.n_owned:].zero_()
        r.dy[:self.n_owned].copy_(self.dy_owned)
        r.backward()
        self.dx_owned = self.pl.exchange.backward(r.dx)
        ex = self.pl.exchange
        if ex.mode == "native" and getattr(ex, "comm", None) is not None:
            ex.comm.allreduce_sum(r.dparams)      # NCCL through the C ABI's own communicator
        else:
            self._allreduce([r.dparams], ex.group)  # one collective over the flat [dphi | dnode] buffer
