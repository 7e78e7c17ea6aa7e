"""torch.autograd bridges over the C ABI -- the Python counterpart of the ChainRules `rrule`s a Julia shim defines
(INTEGRATION.md).  All tensors here are row-major `[items, D]` float32 CUDA buffers; no compute happens in Python.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

Tensor = torch.Tensor

LAUNCHES = {"count": 0}  # kernels launched through libngpde (bench.py reports it as gpu_launches)


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _f32c(t: Optional[Tensor], name: str, device=None) -> Optional[Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.NgpdeError(f"{name} must be a CUDA tensor: the message-passing path has no CPU fallback")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _workspace(nbytes: int, device) -> Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


_FAMILY_FN = {0: "explicit_edge_conv", 1: "vmh_conv", 2: "mppde_conv", 3: "gno_conv"}


def _factored(lib, handle, desc) -> bool:
    """True when this layer call takes the factored GNOConv evaluation (extra GEMM / scale / reduce launches)."""
    cached = getattr(desc, "_ngpde_factored", None)
    if cached is None or cached[0] != handle:
        out = (C.c_int32 * 4)()
        _lib.check(lib.ngpde_conv_kernel_paths(handle, C.byref(desc), out))
        cached = (handle, out[0] == 2)
        try:
            desc._ngpde_factored = cached
        except AttributeError:
            pass
    return cached[1]


class ConvFunction(torch.autograd.Function):
    """y = layer(x; phi_params, node_params) for the four MLP message-passing families."""

    @staticmethod
    def forward(ctx, x: Tensor, phi_params: Tensor, node_params: Optional[Tensor], handle, desc: _lib.ConvDesc,
                snode: Optional[Tensor], edata: Optional[Tensor], theta: Optional[Tensor], dm: int, dy: int):
        lib = _lib.load()
        dev = x.device
        x = _f32c(x, "x")
        phi_params = _f32c(phi_params, "phi parameters")
        node_params = _f32c(node_params, "node parameters")
        N = x.shape[0]
        has_node = desc.node.n_layers > 0
        mbar = torch.empty((N, dm), dtype=torch.float32, device=dev)
        y = torch.empty((N, dy), dtype=torch.float32, device=dev) if has_node else mbar
        with torch.cuda.device(dev):
            nstate = lib.ngpde_conv_state_bytes(handle, C.byref(desc))
        # what the backward would otherwise recompute (hoisted first-layer projections): kept alive with the autograd node
        state = torch.empty(int(nstate), dtype=torch.uint8, device=dev) if nstate else None
        io = _lib.ConvIO(x=_ptr(x), snode=_ptr(snode), edata=_ptr(edata), theta=_ptr(theta),
                         phi_params=_ptr(phi_params), node_params=_ptr(node_params), mbar=_ptr(mbar), y=_ptr(y),
                         state=_ptr(state))
        with torch.cuda.device(dev):
            nbytes = lib.ngpde_conv_workspace_bytes(handle, C.byref(desc), 0)
            if nbytes == 0:
                _lib.check(-1)
            ws = _workspace(nbytes, dev)
            fn = getattr(lib, f"ngpde_{_FAMILY_FN[desc.family]}_forward")
            _lib.check(fn(handle, C.byref(desc), C.byref(io), ws.data_ptr(), ws.numel(), _stream(dev)))
        LAUNCHES["count"] += (2 if has_node else 1) + (1 if _factored(lib, handle, desc) else 0)
        ctx.save_for_backward(x, phi_params, node_params if node_params is not None else x.new_empty(0), mbar)
        ctx.handle, ctx.desc = handle, desc
        ctx.static = (snode, edata, theta)
        ctx.state = state
        ctx.has_node = has_node
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        lib = _lib.load()
        x, phi_params, node_params, mbar = ctx.saved_tensors
        snode, edata, theta = ctx.static
        desc, handle = ctx.desc, ctx.handle
        dev = x.device
        gy = _f32c(gy, "dy")
        dx = torch.empty_like(x)
        dphi = torch.empty_like(phi_params)
        dnode = torch.empty_like(node_params) if ctx.has_node else None
        with torch.cuda.device(dev):
            nbytes = lib.ngpde_conv_workspace_bytes(handle, C.byref(desc), 1)
            if nbytes == 0:
                _lib.check(-1)
            ws = _workspace(nbytes, dev)
            io = _lib.ConvIO(x=_ptr(x), snode=_ptr(snode), edata=_ptr(edata), theta=_ptr(theta),
                             phi_params=_ptr(phi_params), node_params=_ptr(node_params) if ctx.has_node else None,
                             mbar=_ptr(mbar), y=None if not ctx.has_node else _ptr(mbar), dy=_ptr(gy), dx=_ptr(dx),
                             dphi_params=_ptr(dphi), dnode_params=_ptr(dnode), state=_ptr(ctx.state))
            fn = getattr(lib, f"ngpde_{_FAMILY_FN[desc.family]}_backward")
            _lib.check(fn(handle, C.byref(desc), C.byref(io), ws.data_ptr(), ws.numel(), _stream(dev)))
        LAUNCHES["count"] += (8 if ctx.has_node else 5) + (4 if _factored(lib, handle, desc) else 0)
        return dx, dphi, dnode, None, None, None, None, None, None, None


class GcnFunction(torch.autograd.Function):
    """GCNConv forward/backward; params = flat [weight (in*out), bias (out)]."""

    @staticmethod
    def forward(ctx, x: Tensor, params: Tensor, handle, desc: _lib.GcnDesc, edge_weight: Optional[Tensor],
                graph_weight: Optional[Tensor]):
        lib = _lib.load()
        dev = x.device
        x = _f32c(x, "x")
        params = _f32c(params, "parameters")
        edge_weight = _f32c(edge_weight, "edge_weight")
        graph_weight = _f32c(graph_weight, "graph weights")
        N = x.shape[0]
        y = torch.empty((N, desc.out_chs), dtype=torch.float32, device=dev)
        nw = desc.in_chs * desc.out_chs
        with torch.cuda.device(dev):
            nbytes = lib.ngpde_gcn_workspace_bytes(handle, C.byref(desc), 0)
            if nbytes == 0:
                _lib.check(-1)
            ws = _workspace(nbytes, dev)
            bias_ptr = params.data_ptr() + 4 * nw if desc.has_bias else None
            _lib.check(lib.ngpde_gcn_conv_forward(handle, C.byref(desc), x.data_ptr(), params.data_ptr(), bias_ptr,
                                                  _ptr(edge_weight), _ptr(graph_weight), y.data_ptr(), ws.data_ptr(),
                                                  ws.numel(), _stream(dev)))
        LAUNCHES["count"] += 3
        ctx.save_for_backward(x, params, y)
        ctx.handle, ctx.desc, ctx.ew, ctx.gw = handle, desc, edge_weight, graph_weight
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        lib = _lib.load()
        x, params, y = ctx.saved_tensors
        desc, handle = ctx.desc, ctx.handle
        dev = x.device
        gy = _f32c(gy, "dy")
        dx = torch.empty_like(x)
        dparams = torch.zeros_like(params)
        nw = desc.in_chs * desc.out_chs
        with torch.cuda.device(dev):
            nbytes = lib.ngpde_gcn_workspace_bytes(handle, C.byref(desc), 1)
            if nbytes == 0:
                _lib.check(-1)
            ws = _workspace(nbytes, dev)
            bias_ptr = params.data_ptr() + 4 * nw if desc.has_bias else None
            dbias_ptr = dparams.data_ptr() + 4 * nw if desc.has_bias else None
            _lib.check(lib.ngpde_gcn_conv_backward(handle, C.byref(desc), x.data_ptr(), params.data_ptr(), bias_ptr,
                                                   _ptr(ctx.ew), _ptr(ctx.gw), y.data_ptr(), gy.data_ptr(),
                                                   dx.data_ptr(), dparams.data_ptr(), dbias_ptr, ws.data_ptr(),
                                                   ws.numel(), _stream(dev)))
        LAUNCHES["count"] += 8
        return dx, dparams, None, None, None, None


def aggregate(handle, aggr: str, x_rm: Tensor, w: Optional[Tensor] = None) -> Tensor:
    """Bare ordered propagate(copy_xj | e_mul_xj, g, aggr) on a row-major [N, D] tensor (forward only)."""
    lib = _lib.load()
    x_rm = _f32c(x_rm, "x")
    w = _f32c(w, "w")
    out = torch.empty_like(x_rm)
    with torch.cuda.device(x_rm.device):
        _lib.check(lib.ngpde_aggregate(handle, _lib.AGGR[aggr], x_rm.data_ptr(), x_rm.shape[1], _ptr(w), out.data_ptr(),
                                       _stream(x_rm.device)))
    LAUNCHES["count"] += 1
    return out


class WeightedSumFunction(torch.autograd.Function):
    """y = propagate(e_mul_xj, g, +; xj = x, e = w) with a static per-edge weight (SpectralConv, layers.jl:652-657), on the
    ordered aggregate kernel.  The pullback w.r.t. x is the same kernel on the transposed graph (edges kept in their stored
    order, so `w` is shared): dx_j = sum over the out-edges k of j, ascending k, of w_k dy_t(k)."""

    @staticmethod
    def forward(ctx, x_rm: Tensor, g, w: Tensor):
        ctx.g, ctx.w = g, w
        return aggregate(g.handle(x_rm.device), "+", x_rm, w)

    @staticmethod
    def backward(ctx, gy: Tensor):
        gt = ctx.g.transposed()
        return aggregate(gt.handle(gy.device), "+", gy.contiguous(), ctx.w), None, None


def axpy_stages(out: Tensor, u: Optional[Tensor], ks, coefs) -> Tensor:
    """out = u + sum_i coefs[i] * ks[i]  (one fused kernel; ODE stage glue).  u = None stands for zeros; out may alias u."""
    lib = _lib.load()
    nk = len(ks)
    arr = (C.c_void_p * max(nk, 1))(*[k.data_ptr() for k in ks])
    cf = (C.c_float * max(nk, 1))(*[float(c) for c in coefs])
    with torch.cuda.device(out.device):
        _lib.check(lib.ngpde_axpy_stages(out.data_ptr(), _ptr(u), arr, cf, nk, out.numel(), _stream(out.device)))
    LAUNCHES["count"] += 1
    return out


class AxpyStagesFunction(torch.autograd.Function):
    """Differentiable `u + sum_i c_i k_i` (the Runge-Kutta stage combination) on the fused axpy kernel."""

    @staticmethod
    def forward(ctx, coefs, u: Tensor, *ks: Tensor):
        u = u.contiguous()
        ks = [k.contiguous() for k in ks]
        ctx.coefs = tuple(float(c) for c in coefs)
        return axpy_stages(torch.empty_like(u), u, ks, ctx.coefs)

    @staticmethod
    def backward(ctx, g: Tensor):
        g = g.contiguous()
        lib = _lib.load()
        outs = []
        for c in ctx.coefs:
            o = torch.empty_like(g)
            arr = (C.c_void_p * 1)(g.data_ptr())
            cf = (C.c_float * 1)(c)
            with torch.cuda.device(g.device):
                _lib.check(lib.ngpde_axpy_stages(o.data_ptr(), None, arr, cf, 1, g.numel(), _stream(g.device)))
            LAUNCHES["count"] += 1
            outs.append(o)
        return (None, g, *outs)
