"""Optimisers over the flat parameter vector, on the device.

Mirrors the slice of Optimisers.jl the reference's tutorials use ([DEP], Project.toml; call sites
/root/reference/docs/src/tutorials/graph_node.md:122-129 `Optimisers.Adam(0.01f0)` and VMH.md:97,140-143
`Rprop(1.0f-6, (5.0f-1, 1.2f0), (1.0f-8, 10.0f0))`):

    st_opt = setup(opt, ps)                      # ps: ComponentArray (or a flat float32 CUDA tensor)
    st_opt, ps = update(st_opt, ps, gs)          # in place on the flat buffer, one fused kernel (ngpde_adam_step / ngpde_rprop_step)

Update rules (Optimisers.jl `apply!`; restated for the parity tests in oracle/ngpde_oracle.py):
    Adam   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  x -= m / (1-b1^t) / (sqrt(v / (1-b2^t)) + eps) * eta
    Rprop  eta_i <- min(eta_i l+, G+) if g_prev g > 0, max(eta_i l-, G-) if < 0;  g_prev <- 0 if g_prev g < 0 else g;
           x -= eta_i sign(g_prev)
There is no CPU fallback: the state lives where the parameters live, and they live on the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np
import torch

from . import _lib, ops
from .lux import ComponentArray

Tensor = torch.Tensor


def _flat(ps) -> Tensor:
    t = ps.data if isinstance(ps, ComponentArray) else ps
    if not isinstance(t, Tensor) or not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
        raise _lib.NgpdeError("optimiser: parameters must be a contiguous float32 CUDA vector (ComponentArray(ps))")
    return t.detach()


@dataclass
class Adam:
    eta: float = 0.001
    beta: Tuple[float, float] = (0.9, 0.999)
    epsilon: float = 1e-8


@dataclass
class Rprop:
    eta: float = 1e-3
    ell: Tuple[float, float] = (0.5, 1.2)
    gamma: Tuple[float, float] = (1e-6, 50.0)


class _AdamState:
    def __init__(self, opt: Adam, x: Tensor):
        self.opt = opt
        self.m, self.v = torch.zeros_like(x), torch.zeros_like(x)
        # Optimisers.jl carries beta^t as Float32 pairs: init = beta, then `bt .* beta` after every step
        self.bt = (np.float32(opt.beta[0]), np.float32(opt.beta[1]))


class _RpropState:
    def __init__(self, opt: Rprop, x: Tensor):
        self.opt = opt
        self.g = torch.zeros_like(x)
        self.eta = torch.full_like(x, float(np.float32(opt.eta)))


def setup(opt, ps):
    """Optimisers.setup(opt, ps)."""
    x = _flat(ps)
    if isinstance(opt, Adam):
        return _AdamState(opt, x)
    if isinstance(opt, Rprop):
        return _RpropState(opt, x)
    raise TypeError(f"unsupported optimiser {type(opt).__name__}; Adam and Rprop are the ones the reference's tutorials use")


def update(state, ps, gs):
    """Optimisers.update(st_opt, ps, gs): one fused kernel over the flat vector, in place; returns (state, ps)."""
    x, g = _flat(ps), _flat(gs)
    if g.numel() != x.numel():
        raise ValueError(f"gradient has {g.numel()} entries, parameters {x.numel()}")
    lib = _lib.load()
    f32 = lambda v: float(np.float32(v))
    with torch.cuda.device(x.device):
        st = ops._stream(x.device)
        if isinstance(state, _AdamState):
            o = state.opt
            _lib.check(lib.ngpde_adam_step(x.data_ptr(), g.data_ptr(), state.m.data_ptr(), state.v.data_ptr(), x.numel(),
                                           f32(o.eta), f32(o.beta[0]), f32(o.beta[1]), f32(o.epsilon), float(state.bt[0]),
                                           float(state.bt[1]), st))
            state.bt = (np.float32(state.bt[0] * np.float32(o.beta[0])), np.float32(state.bt[1] * np.float32(o.beta[1])))
        elif isinstance(state, _RpropState):
            o = state.opt
            _lib.check(lib.ngpde_rprop_step(x.data_ptr(), g.data_ptr(), state.g.data_ptr(), state.eta.data_ptr(), x.numel(),
                                            f32(o.ell[0]), f32(o.ell[1]), f32(o.gamma[0]), f32(o.gamma[1]), st))
        else:
            raise TypeError("update: unknown optimiser state")
    ops.LAUNCHES["count"] += 1
    return state, ps
