"""Fixed-step Neural-ODE trajectory parity (north_star: <= 1e-4 on the trajectory at fixed step), forward and the
discrete-adjoint gradient, for the C1 (Tsit5) and C3 (RK4) model shapes."""
import numpy as np
import pytest
import torch

import ngpde
import ngpde_oracle as orc
from common import flat_grad, oracle_forward, relerr, to_ograph, tree_requires_grad, tree_to_cpu
from ngpde import ode, workloads

pytestmark = pytest.mark.gpu
TRAJ_TOL = 1e-4


def _oracle_traj(w, method, dt, t1, dtype, with_grad):
    og = to_ograph(w.graph, dtype)
    pc = tree_to_cpu(w.ps, dtype)
    if with_grad:
        tree_requires_grad(pc)
    u0 = w.x.detach().cpu().to(dtype)
    uT = orc.integrate_fixed(lambda u: oracle_forward(w.layer, u, pc, og), u0, 0.0, t1, dt, method)
    if not with_grad:
        return uT.detach(), None
    (uT ** 2).mean().backward()
    return uT.detach(), flat_grad(pc)


@pytest.mark.parametrize("name,method,kw,dt,t1", [("c1", "tsit5", {}, 0.05, 1.0), ("c3", "rk4", {"side": 20}, 0.05, 0.5)])
def test_trajectory_and_adjoint(name, method, kw, dt, t1):
    w = workloads.WORKLOADS[name]("cuda", **kw)
    ca = ngpde.ComponentArray(w.ps)
    ca.data.requires_grad_(True)
    uT, _, evals = ode.solve_fixed(w.layer, w.x, ca, w.st, (0.0, t1), dt, method)
    nsteps = int(round(t1 / dt))
    assert evals == nsteps * (6 if method == "tsit5" else 4)
    (uT ** 2).mean().backward()
    u32, g32 = _oracle_traj(w, method, dt, t1, torch.float32, True)
    u64, g64 = _oracle_traj(w, method, dt, t1, torch.float64, True)
    errs = dict(u32=relerr(uT, u32), u64=relerr(uT, u64), g32=relerr(ca.data.grad, g32), g64=relerr(ca.data.grad, g64))
    assert all(v <= TRAJ_TOL for v in errs.values()), errs


@pytest.mark.parametrize("name,method,kw,dt,t1", [("c1", "tsit5", {}, 0.05, 1.0), ("c3", "rk4", {"side": 20}, 0.05, 0.5)])
def test_cuda_graph_step_and_adjoint(name, method, kw, dt, t1):
    """The CUDA-graph-captured RK step and its captured discrete adjoint (ode.GraphedRK: only libngpde kernels inside)
    against the oracle's trajectory / autograd gradient, and bit-for-bit repeatable."""
    w = workloads.WORKLOADS[name]("cuda", **kw)
    nsteps = int(round(t1 / dt))
    rk = ode.GraphedRK(w.layer, w.x, w.ps, w.st, dt, method)
    assert rk.kernels_fwd and rk.kernels_bwd  # both graphs hold kernel nodes only from this library
    uT = rk.solve(w.x, nsteps).clone()
    n = uT.numel()
    lam, dps = rk.adjoint(2.0 * uT / n)   # L = mean(u_T^2)
    u64, g64 = _oracle_traj(w, method, dt, t1, torch.float64, True)
    r0 = rk.stage[0]

    def unpad(flat):  # [dphi | pad | dnode] -> ComponentArray order
        parts = [flat[:r0.phi.numel()]]
        if r0.node is not None:
            parts.append(flat[flat.numel() - r0.node.numel():])
        return torch.cat(parts)

    g = unpad(dps)
    errs = dict(u=relerr(uT.T, u64), g=relerr(g, g64))
    assert all(v <= TRAJ_TOL for v in errs.values()), errs
    # same trajectory through the eager autograd path (every RHS is the same kernel): the two must agree closely
    ca = ngpde.ComponentArray(w.ps)
    ca.data.requires_grad_(True)
    uT2, _, _ = ode.solve_fixed(w.layer, w.x, ca, w.st, (0.0, t1), dt, method)
    (uT2 ** 2).mean().backward()
    assert relerr(uT.T, uT2) <= 1e-6 and relerr(g, ca.data.grad) <= 1e-5
    # replays are deterministic
    uT3 = rk.solve(w.x, nsteps).clone()
    _, dps3 = rk.adjoint(2.0 * uT3 / n)
    assert torch.equal(uT, uT3) and torch.equal(g, unpad(dps3))


@pytest.mark.parametrize("method,dt,t1,side,hidden", [("tsit5", 0.05, 1.0, 32, 16),    # C1: tensor-core kernels, 16-CTA cluster
                                                        ("rk4", 0.05, 0.5, 17, 16),     # ragged node ranges
                                                        ("rk4", 0.05, 0.25, 9, 12),     # 81 nodes: the portable 8-CTA cluster, padded widths
                                                        ("tsit5", 0.05, 0.25, 12, 24)])  # layers 17..32 wide: the FFMA kernels
def test_persistent_kernel_integrator_and_adjoint(method, dt, t1, side, hidden):
    """One-launch integrator + one-launch adjoint (ode.PersistentRK, thread-block cluster) on the C1 model: trajectory and
    gradient against the float64 oracle (<= 1e-4) and against the CUDA-graph path built from the layer kernels."""
    w = workloads.c1_edgeconv("cuda", side=side, hidden=hidden)
    nsteps = int(round(t1 / dt))
    prk = ode.PersistentRK(w.layer, w.x, w.ps, w.st, dt, method)
    uT = prk.solve(w.x, nsteps).clone()
    n = uT.numel()
    lam, dps = prk.adjoint(2.0 * uT / n)
    u64, g64 = _oracle_traj(w, method, dt, t1, torch.float64, True)
    errs = dict(u=relerr(uT.T, u64), g=relerr(dps, g64))
    assert all(v <= TRAJ_TOL for v in errs.values()), errs
    rk = ode.GraphedRK(w.layer, w.x, w.ps, w.st, dt, method)
    uT2 = rk.solve(w.x, nsteps).clone()
    lam2, dps2 = rk.adjoint(2.0 * uT2 / n)
    assert relerr(uT, uT2) <= 1e-6 and relerr(lam, lam2) <= 1e-5 and relerr(dps, dps2[:dps.numel()]) <= 1e-5
    # deterministic
    uT3 = prk.solve(w.x, nsteps).clone()
    _, dps3 = prk.adjoint(2.0 * uT3 / n)
    assert torch.equal(uT, uT3) and torch.equal(dps, dps3)


def test_persistent_kernel_rejects_what_it_cannot_take():
    w = workloads.c3_vmh("cuda", side=8)
    with pytest.raises(TypeError):
        ode.PersistentRK(w.layer, w.x, w.ps, w.st, 0.1)
