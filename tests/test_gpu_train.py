"""Loss + optimiser kernels against the oracle's float32 restatement of Optimisers.jl / the
tutorials' loss functions.  Element-wise updates: bit-exact.  Reduced scalars and cotangents: <= 1e-6 relative."""
import numpy as np
import pytest
import torch

import ngpde
import ngpde_oracle as orc
from common import relerr
from ngpde import losses, optim

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("n", [1, 25282, 1_000_003])
def test_adam_bit_exact(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32)
    xd = torch.from_numpy(x.copy()).to(DEV)
    opt = optim.Adam(0.01)  # graph_node.md:122
    st = optim.setup(opt, xd)
    so = orc.adam_init(x, opt.beta)
    for it in range(5):
        g = (rng.standard_normal(n) * 10.0 ** rng.integers(-6, 2)).astype(np.float32)
        st, _ = optim.update(st, xd, torch.from_numpy(g).to(DEV))
        x, so = orc.adam_step(x, g, so, opt.eta, opt.beta, opt.epsilon)
        assert np.array_equal(xd.cpu().numpy(), x), f"iteration {it}"
        assert np.array_equal(st.m.cpu().numpy(), so[0]) and np.array_equal(st.v.cpu().numpy(), so[1])


@pytest.mark.parametrize("n", [7, 25282])
def test_rprop_bit_exact(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32)
    xd = torch.from_numpy(x.copy()).to(DEV)
    opt = optim.Rprop(1.0e-6, (5.0e-1, 1.2), (1.0e-8, 10.0))  # VMH.md:97
    st = optim.setup(opt, xd)
    so = orc.rprop_init(x, opt.eta)
    for it in range(8):
        g = rng.standard_normal(n).astype(np.float32)
        g[rng.integers(0, n, max(1, n // 5))] = 0.0  # exercise the g*dx == 0 branch
        st, _ = optim.update(st, xd, torch.from_numpy(g).to(DEV))
        x, so = orc.rprop_step(x, g, so, opt.ell, opt.gamma)
        assert np.array_equal(xd.cpu().numpy(), x), f"iteration {it}"
        assert np.array_equal(st.eta.cpu().numpy(), so[1]) and np.array_equal(st.g.cpu().numpy(), so[0])


def test_optimiser_accepts_component_array_and_rejects_cpu():
    l = ngpde.Dense(3, 4)
    ps, _ = ngpde.setup(0, l, DEV)
    ca = ngpde.ComponentArray(ps)
    st = optim.setup(optim.Adam(), ca)
    before = ca.data.clone()
    optim.update(st, ca, torch.ones_like(ca.data))
    assert not torch.equal(before, ca.data)
    with pytest.raises(ngpde.NgpdeError):
        optim.setup(optim.Adam(), torch.zeros(4))


@pytest.mark.parametrize("shape", [(1, 5), (2, 65536), (3, 700_001)])
def test_mse_loss_and_cotangent(shape):
    rng = np.random.default_rng(1)
    yh = torch.from_numpy(rng.standard_normal(shape[::-1]).astype(np.float32)).T  # Julia-shaped, column-major
    y = torch.from_numpy(rng.standard_normal(shape[::-1]).astype(np.float32)).T
    a = yh.to(DEV).requires_grad_(True)
    l = losses.mse(a, y.to(DEV))
    l.backward()
    b = yh.double().requires_grad_(True)
    lo = orc.mse(b, y.double())
    lo.backward()
    assert abs(l.item() - lo.item()) <= 1e-6 * abs(lo.item())
    assert relerr(a.grad, b.grad) <= 1e-6
    l2 = losses.mse(yh.to(DEV), y.to(DEV))
    assert l2.item() == l.item()  # deterministic reduction


@pytest.mark.parametrize("n,c,masked", [(2708, 7, True), (50, 3, False), (100_000, 10, True)])
def test_logitcrossentropy_loss_and_cotangent(n, c, masked):
    rng = np.random.default_rng(2)
    yh = torch.from_numpy((3 * rng.standard_normal((n, c))).astype(np.float32)).T
    mask = torch.from_numpy(np.sort(rng.choice(n, max(1, n // 10), replace=False))) if masked else None
    nm = n if mask is None else mask.numel()
    y = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, c, nm)), c).float().T  # (c, nm) one-hot
    a = yh.to(DEV).requires_grad_(True)
    l = losses.logitcrossentropy(a, y.to(DEV), None if mask is None else mask.to(DEV))
    l.backward()
    b = yh.double().requires_grad_(True)
    lo = orc.logitcrossentropy(b[:, mask] if mask is not None else b, y.double())  # graph_node.md:104
    lo.backward()
    assert abs(l.item() - lo.item()) <= 2e-6 * abs(lo.item())
    assert relerr(a.grad, b.grad) <= 2e-6
    if mask is not None:
        other = torch.ones(n, dtype=torch.bool)
        other[mask] = False
        assert torch.count_nonzero(a.grad.cpu()[:, other]) == 0


def test_training_iteration_stays_on_device():
    """The graph_node.md:122-129 loop -- layer forward -> masked cross-entropy -> pullback -> Adam -- with every step on
    libngpde kernels, against the same loop run by the oracle on the CPU (oracle layers, torch autograd, oracle Adam):
    the loss trajectory and the final parameters must agree."""
    from common import oracle_forward, to_ograph, tree_requires_grad, tree_to_cpu, flat_grad, tree_leaves
    rng = np.random.default_rng(3)
    n, e, c = 300, 2400, 4
    s, t = rng.integers(0, n, e), rng.integers(0, n, e)
    g = ngpde.GNNGraph(torch.from_numpy(s), torch.from_numpy(t), num_nodes=n).to(DEV)
    model = ngpde.Chain(ngpde.GCNConv((8, 16), "relu", initialgraph=g), ngpde.GCNConv((16, c), initialgraph=g))
    ps, st = ngpde.setup(rng, model, DEV)
    ca = ngpde.ComponentArray(ps)
    x = torch.from_numpy(rng.standard_normal((n, 8)).astype(np.float32)).T
    mask = torch.arange(0, n, 3)
    y = torch.nn.functional.one_hot(torch.from_numpy(rng.integers(0, c, mask.numel())), c).float().T
    # ---- oracle loop ----
    pc = tree_requires_grad(tree_to_cpu(ps))
    og = to_ograph(g)
    flat = lambda: torch.cat([v.detach().T.reshape(-1) if v.dim() == 2 else v.detach().reshape(-1) for v in tree_leaves(pc)])
    so = orc.adam_init(flat().numpy(), (0.9, 0.999))
    hist_o = []
    for _ in range(12):
        for v in tree_leaves(pc):
            v.grad = None
        l = orc.logitcrossentropy(oracle_forward(model, x, pc, og)[:, mask], y)
        l.backward()
        hist_o.append(l.item())
        new, so = orc.adam_step(flat().numpy(), flat_grad(pc).numpy(), so, 0.01)
        off = 0
        with torch.no_grad():
            for v in tree_leaves(pc):
                k = v.numel()
                blk = torch.from_numpy(new[off:off + k])
                v.copy_(blk.reshape(v.shape[::-1]).T if v.dim() == 2 else blk.reshape(v.shape))
                off += k
    # ---- product loop ----
    ca.data.requires_grad_(True)
    xd, yd, md = x.to(DEV), y.to(DEV), mask.to(DEV)
    st_opt = optim.setup(optim.Adam(0.01), ca)
    hist = []
    for _ in range(12):
        ca.data.grad = None
        yh, _ = model(xd, ca, st)
        l = losses.logitcrossentropy(yh, yd, md)
        l.backward()
        st_opt, _ = optim.update(st_opt, ca, ca.data.grad)
        hist.append(l.item())
    assert hist[-1] < hist[0]
    assert max(abs(a - b) / abs(b) for a, b in zip(hist, hist_o)) <= 1e-5, (hist, hist_o)
    assert relerr(ca.data.detach(), torch.from_numpy(new)) <= 1e-4
