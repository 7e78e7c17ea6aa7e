"""Fixed-step explicit Runge-Kutta drivers around a layer call -- the immediate caller of the hot path
(`dudt(u, p, t) = model(u, p, st)[1]`, /root/reference/docs/src/tutorials/graph_node.md:59-66, VMH.md:87; the solvers
themselves are DifferentialEquations.jl [DEP]: `RK4()` and `Tsit5()` with `adaptive=false, dt=...`).

Every right-hand-side evaluation is one fused layer call; the stage combinations `u + dt * sum a_ij k_j` run on the
library's axpy kernel (ngpde_axpy_stages).  Differentiating through `solve_fixed` with torch.autograd is the
discrete adjoint of the fixed-step scheme: each stage's pullback is the layer's hand-written backward kernel.
Node states are kept as row-major `[N, d]` buffers between stages (`x.T` views for the Lux-shaped layer call).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch

from .ops import AxpyStagesFunction

Tensor = torch.Tensor

# Butcher tableaus: (c, a rows, b)
RK4 = (
    (0.0, 0.5, 0.5, 1.0),
    ((), (0.5,), (0.0, 0.5), (0.0, 0.0, 1.0)),
    (1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0),
)

# Tsitouras 5(4) (Tsitouras 2011, as used by OrdinaryDiffEq.Tsit5); the 7th stage is FSAL and equals the next step's
# first stage, so a fixed-step run costs 6 new right-hand sides per step.
_T5A = (
    (),
    (0.161,),
    (-0.008480655492356989, 0.335480655492357),
    (2.8971530571054935, -6.359448489975075, 4.3622954328695815),
    (5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525),
    (5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383),
)
_T5B = (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774)
TSIT5 = ((0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0), _T5A, _T5B)

TABLEAUS = {"rk4": RK4, "tsit5": TSIT5}


def _combine(u: Tensor, ks: Sequence[Tensor], coefs: Sequence[float]) -> Tensor:
    nz = [(c, k) for c, k in zip(coefs, ks) if c != 0.0]
    if not nz:
        return u
    return AxpyStagesFunction.apply(tuple(c for c, _ in nz), u, *[k for _, k in nz])


def rk_step(rhs: Callable[[Tensor], Tensor], u: Tensor, dt: float, tableau=RK4, k1: Optional[Tensor] = None):
    """One explicit RK step on a row-major state.  Returns (u_next, n_rhs_evaluated)."""
    _, A, b = tableau
    ks: List[Tensor] = []
    n = 0
    for i, row in enumerate(A):
        if i == 0 and k1 is not None:
            ks.append(k1)
            continue
        ui = _combine(u, ks, [dt * a for a in row])
        ks.append(rhs(ui))
        n += 1
    return _combine(u, ks, [dt * w for w in b]), n


def solve_fixed(layer, x: Tensor, ps, st, tspan: Tuple[float, float], dt: float, method: str = "rk4",
                saveat_every: int = 0):
    """Integrate du/dt = layer(u, ps, st)[1] from tspan[0] to tspan[1] with a fixed step.

    x is Lux-shaped (d, N); returns (u_final (d, N), [saved states], number of RHS evaluations)."""
    tab = TABLEAUS[method]
    nsteps = int(round((tspan[1] - tspan[0]) / dt))
    u = x.T.contiguous()  # [N, d] row-major
    evals = 0

    def rhs(u_rm: Tensor) -> Tensor:
        y, _ = layer(u_rm.T, ps, st)
        return y.T

    saved = []
    for step in range(nsteps):
        u, n = rk_step(rhs, u, dt, tab)
        evals += n
        if saveat_every and (step + 1) % saveat_every == 0:
            saved.append(u.T)
    return u.T, saved, evals
