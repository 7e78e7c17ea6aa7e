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


class GraphedRK:
    """One explicit Runge-Kutta step of `du/dt = layer(u, ps, st)[1]` and its discrete adjoint, each captured ONCE as a
    CUDA graph and replayed per step (the reference's solver loop is `solve(prob, Tsit5() | RK4();
    adaptive = false, dt)`, docs/src/tutorials/graph_node.md:53-66, VMH.md:87).

    Forward graph (S = number of stages):   for s: u_s = u + dt sum_{j<s} a_sj k_j   (ngpde_axpy_stages)
                                                   k_s = layer(u_s)                  (fused layer forward, C ABI)
                                            u <- u + dt sum_s b_s k_s
    Adjoint graph, given lam = dL/du_next:  for s = S..1: kbar_s = dt b_s lam + dt sum_{i>s} a_is ubar_i
                                                          (ubar_s, dps_s) = layer VJP at u_s with cotangent kbar_s
                                            lam <- lam + sum_s ubar_s;   dparams += sum_s dps_s
    Both graphs contain only libngpde kernels: no Python, no allocator and no autograd bookkeeping per RHS.  The stage
    states u_s live in the stage runners (re-established by replaying the forward graph from the step's checkpointed u), so
    a trajectory costs O(steps) state copies, and its gradient two forward replays + one adjoint replay per step.
    """

    def __init__(self, layer, x: Tensor, ps, st, dt: float, method: str = "rk4"):
        from .engine import RhsRunner
        self.c, self.A, self.b = TABLEAUS[method]
        self.dt, self.S = float(dt), len(self.A)
        first = RhsRunner(layer, x, ps, st)
        self.stage: List[RhsRunner] = [first] + [RhsRunner(layer, x, ps, st, share=first) for _ in range(self.S - 1)]
        self.dev = first.dev
        self.u = first.x.clone()              # [N, d] current state (row-major image of Julia's (d, N))
        self.lam = torch.zeros_like(self.u)   # adjoint state dL/du
        self.dparams = torch.zeros_like(first.dparams)
        self._tmp = torch.empty_like(self.u)
        self.fwd_graph = self.bwd_graph = None
        self.kernels_fwd = self.kernels_bwd = None
        self.rhs_evals = 0
        if first.y.shape != self.u.shape:
            raise ValueError("an ODE right-hand side must map the state onto its own shape")
        self._capture()

    # ---- the two step bodies (launch-only: safe to capture) ----
    def _fwd_body(self):
        from .ops import axpy_stages
        dt = self.dt
        for s_, r in enumerate(self.stage):
            nz = [(dt * a, self.stage[j].y) for j, a in enumerate(self.A[s_]) if a != 0.0]
            axpy_stages(r.x, self.u, [k for _, k in nz], [c for c, _ in nz])
            r.forward()
        nz = [(dt * w, self.stage[j].y) for j, w in enumerate(self.b) if w != 0.0]
        axpy_stages(self.u, self.u, [k for _, k in nz], [c for c, _ in nz])

    def _bwd_body(self):
        from .ops import axpy_stages
        dt = self.dt
        for s_ in range(self.S - 1, -1, -1):
            r = self.stage[s_]
            ks, cs = [self.lam], [dt * self.b[s_]]
            for i in range(s_ + 1, self.S):
                a = self.A[i][s_] if s_ < len(self.A[i]) else 0.0
                if a != 0.0:
                    ks.append(self.stage[i].dx)
                    cs.append(dt * a)
            axpy_stages(r.dy, None, ks, cs)   # kbar_s (u = NULL stands for zeros)
            r.backward()                      # -> r.dx = ubar_s, r.dparams = dps_s
        axpy_stages(self.lam, self.lam, [r.dx for r in self.stage], [1.0] * self.S)
        axpy_stages(self.dparams, self.dparams, [r.dparams for r in self.stage], [1.0] * self.S)

    def _capture(self):
        from . import _lib
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        u0 = self.u.clone()
        with torch.cuda.stream(side):   # warm-up outside capture (lazy layout builds, function attributes)
            self._fwd_body()
            self._bwd_body()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        graphs = []
        for body in (self._fwd_body, self._bwd_body):
            g = torch.cuda.CUDAGraph(keep_graph=True)
            with torch.cuda.graph(g):
                body()
            try:
                nk, _ = _lib.cuda_graph_kernel_nodes(g.raw_cuda_graph())
            except Exception:  # noqa: BLE001
                nk = None
            g.instantiate()
            graphs.append((g, nk))
        (self.fwd_graph, self.kernels_fwd), (self.bwd_graph, self.kernels_bwd) = graphs
        self.u.copy_(u0)
        self.lam.zero_()
        self.dparams.zero_()

    # ---- public ----
    def set_params(self, ps) -> None:
        self.stage[0].set_params(ps)  # the other stages share its parameter buffers

    def step(self) -> Tensor:
        """u <- RK step(u): one graph replay."""
        self.fwd_graph.replay()
        self.rhs_evals += self.S
        return self.u

    def solve(self, u0: Tensor, nsteps: int, keep: bool = True) -> Tensor:
        """Integrate `nsteps` steps from u0 (Julia-shaped (d, N) or row-major [N, d]); keeps the per-step states for `adjoint`."""
        u0 = u0.T if u0.shape != self.u.shape else u0
        self.u.copy_(u0)
        self._ckpt = [] if keep else None
        for _ in range(nsteps):
            if keep:
                self._ckpt.append(self.u.clone())
            self.step()
        return self.u

    def adjoint(self, dL_duT: Tensor):
        """Discrete adjoint of the last `solve`: returns (dL/du0 [N, d], dL/dparams flat [dphi | pad | dnode])."""
        g = dL_duT.T if dL_duT.shape != self.u.shape else dL_duT
        self.lam.copy_(g)
        self.dparams.zero_()
        for u_n in reversed(self._ckpt):
            self.u.copy_(u_n)
            self.fwd_graph.replay()   # re-establish the stage states u_s, k_s of this step
            self.bwd_graph.replay()
            self.rhs_evals += self.S
        return self.lam, self.dparams


class PersistentRK:
    """The whole fixed-step integration of `du/dt = ExplicitEdgeConv(u, ps, st)[1]` in ONE kernel launch, and its discrete
    adjoint in a second one (`ngpde_edgeconv_ode_forward/adjoint`, csrc/ngpde_ode.cu: a thread-block cluster owns the graph,
    one cluster barrier per right-hand side).  For graphs where a right-hand side is launch latency rather than work --
    BASELINE config C1: 1,024 nodes, 121 RHS per trajectory.  Same results as `GraphedRK` / `solve_fixed` up to float32
    rounding of the stage combinations (the aggregation order is the layer kernels').

        rk = PersistentRK(layer, x, ps, st, dt, "tsit5")
        uT = rk.solve(x, nsteps)                  # [N, d]
        du0, dps = rk.adjoint(dL_duT)             # [N, d], flat phi-parameter gradient
    """

    def __init__(self, layer, x: Tensor, ps, st, dt: float, method: str = "rk4"):
        import ctypes as C
        from . import _lib
        from .layers import ExplicitEdgeConv
        if not isinstance(layer, ExplicitEdgeConv):
            raise TypeError("PersistentRK integrates ExplicitEdgeConv right-hand sides; use GraphedRK for the other layers")
        self._lib = _lib
        self.lib = _lib.load()
        (x_rm, phi, _, self.handle, self.desc, self.snode, _, _, _, _) = layer.prepare(x, ps, st)
        self._layer, self._st = layer, st
        self.dev = x_rm.device
        self.phi = phi.detach().contiguous().clone()
        _, A, b = TABLEAUS[method]
        tab = _lib.RkTableau()
        tab.n_stages = len(A)
        for i, row in enumerate(A):
            for j, a in enumerate(row):
                tab.a[i][j] = float(a)
        for i, w in enumerate(b):
            tab.b[i] = float(w)
        self.tab, self.S, self.dt = tab, len(A), float(dt)
        with torch.cuda.device(self.dev):
            nbytes = self.lib.ngpde_edgeconv_ode_workspace_bytes(self.handle, C.byref(self.desc), C.byref(tab))
        if nbytes == 0:
            _lib.check(-1)
        self.ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.dev)
        self.u = x_rm.detach().clone()
        self.lam = torch.zeros_like(self.u)
        self.dparams = torch.zeros_like(self.phi)
        self.traj: Optional[Tensor] = None
        self.nsteps = 0
        self.rhs_evals = 0

    def set_params(self, ps) -> None:
        pr = self._layer.prepare(self.u.T, ps, self._st)
        self.phi.copy_(pr[1].detach())

    def solve(self, u0: Tensor, nsteps: int) -> Tensor:
        import ctypes as C
        from .ops import _ptr, _stream, LAUNCHES
        u0 = u0.T if u0.shape != self.u.shape else u0
        self.u.copy_(u0)
        if self.traj is None or self.nsteps != nsteps:
            self.traj = torch.empty((nsteps, self.S) + tuple(self.u.shape), dtype=torch.float32, device=self.dev)
            self.nsteps = nsteps
        with torch.cuda.device(self.dev):
            self._lib.check(self.lib.ngpde_edgeconv_ode_forward(
                self.handle, C.byref(self.desc), C.byref(self.tab), self.dt, nsteps, self.phi.data_ptr(), _ptr(self.snode),
                self.u.data_ptr(), self.traj.data_ptr(), self.ws.data_ptr(), self.ws.numel(), _stream(self.dev)))
        LAUNCHES["count"] += 1
        self.rhs_evals += nsteps * self.S
        return self.u

    def adjoint(self, dL_duT: Tensor):
        import ctypes as C
        from .ops import _ptr, _stream, LAUNCHES
        g = dL_duT.T if dL_duT.shape != self.u.shape else dL_duT
        self.lam.copy_(g)
        with torch.cuda.device(self.dev):
            self._lib.check(self.lib.ngpde_edgeconv_ode_adjoint(
                self.handle, C.byref(self.desc), C.byref(self.tab), self.dt, self.nsteps, self.phi.data_ptr(), _ptr(self.snode),
                self.traj.data_ptr(), self.lam.data_ptr(), self.dparams.data_ptr(), self.ws.data_ptr(), self.ws.numel(),
                _stream(self.dev)))
        LAUNCHES["count"] += 1
        self.rhs_evals += self.nsteps * self.S
        return self.lam, self.dparams
