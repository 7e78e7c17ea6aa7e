"""Pins the CPU oracle against everything numerical the reference's own tests hold for this path
(/root/reference/test/runtests.jl:153-162 and the doctest /root/reference/src/layers.jl:581-631), and checks its
index/layout helpers against brute force.  CPU only."""
import math

import numpy as np
import torch

import ngpde_oracle as orc


def _x(n, dtype):
    return torch.from_numpy(np.linspace(0.0, 2.0 * math.pi, n + 1)[1:]).to(dtype).reshape(1, -1)


def test_spectralconv_kat_float32():
    # runtests.jl:159-161: sum(abs2, s(sin.(x)) .- cos.(x)) < 1f-3 and the cos/-sin twin, Float32
    g = orc.spectral_graph(100, torch.float32)
    x = _x(100, torch.float32)
    assert ((orc.spectral_conv(torch.sin(x), g, 100) - torch.cos(x)) ** 2).sum().item() < 1e-3
    assert ((orc.spectral_conv(torch.cos(x), g, 100) + torch.sin(x)) ** 2).sum().item() < 1e-3


def test_spectralconv_doctest_float64():
    # layers.jl:591-630: Float64 residuals of the spectral derivative are 1e-16 .. 4e-13
    g = orc.spectral_graph(100, torch.float64)
    x = _x(100, torch.float64)
    assert (orc.spectral_conv(torch.sin(x), g, 100) - torch.cos(x)).abs().max().item() < 1e-11
    assert (orc.spectral_conv(torch.cos(x), g, 100) + torch.sin(x)).abs().max().item() < 1e-11


def test_direction_is_source_to_target():
    # flipping the direction breaks the known answer by O(100)
    g = orc.spectral_graph(100, torch.float64)
    x = _x(100, torch.float64)
    flipped = orc.OGraph(g.t, g.s, g.num_nodes, 1, {}, g.edata, {})
    assert ((orc.spectral_conv(torch.sin(x), flipped, 100) - torch.cos(x)) ** 2).sum().item() > 1.0


def test_ordered_scatter_matches_sequential_loop():
    rng = np.random.default_rng(0)
    n, e, d = 17, 300, 5
    idx = rng.integers(0, n - 2, e)  # last two nodes isolated
    src = torch.from_numpy(rng.standard_normal((d, e)).astype(np.float32))
    for op in ("+", "mean", "max", "min", "*"):
        got = orc.scatter(op, src, idx, n)
        ident = {"+": 0.0, "mean": 0.0, "max": -math.inf, "min": math.inf, "*": 1.0}[op]
        ref = np.full((d, n), ident, dtype=np.float32)
        cnt = np.zeros(n)
        for k in range(e):  # NNlib.scatter's CPU loop
            j = idx[k]
            v = src[:, k].numpy()
            if op in ("+", "mean"):
                ref[:, j] = ref[:, j] + v
            elif op == "max":
                ref[:, j] = np.maximum(ref[:, j], v)
            elif op == "min":
                ref[:, j] = np.minimum(ref[:, j], v)
            else:
                ref[:, j] = ref[:, j] * v
            cnt[j] += 1
        if op == "mean":
            ref = np.where(cnt > 0, ref / np.maximum(cnt, 1).astype(np.float32), 0).astype(np.float32)
        assert np.array_equal(got.numpy(), ref), op  # bit-exact
        assert got.shape == (d, n)


def test_scatter_pullbacks_match_autograd_of_dense_formulation():
    rng = np.random.default_rng(1)
    n, e, d = 6, 20, 3
    idx = rng.integers(0, n, e)
    src = torch.from_numpy(rng.standard_normal((d, e))).requires_grad_(True)
    onehot = torch.zeros(e, n, dtype=torch.float64)
    onehot[torch.arange(e), torch.from_numpy(idx)] = 1
    w = torch.from_numpy(rng.standard_normal((d, n)))
    for op in ("+", "mean"):
        (g1,) = torch.autograd.grad((orc.scatter(op, src, idx, n) * w).sum(), src)
        dense = src @ onehot
        if op == "mean":
            dense = dense / onehot.sum(0).clamp(min=1)
        (g2,) = torch.autograd.grad((dense * w).sum(), src)
        assert torch.allclose(g1, g2, atol=1e-12)


def test_prod_scatter_pullback_is_product_of_others():
    # [DEP] NNlib's pullback of scatter(*, src, idx): gather(dy, idx)[:, k] * prod of the OTHER sources of idx[k], ascending
    # (a left fold) -- checked against the literal loop, with an exact zero among the sources (no division by src)
    rng = np.random.default_rng(2)
    n, e, d = 7, 30, 2
    idx = rng.integers(0, n - 1, e)
    src_np = rng.standard_normal((d, e)).astype(np.float32)
    src_np[:, 4] = 0.0
    src = torch.from_numpy(src_np).requires_grad_(True)
    dy = torch.from_numpy(rng.standard_normal((d, n)).astype(np.float32))
    (got,) = torch.autograd.grad(orc.scatter("*", src, idx, n), src, dy)
    ref = np.zeros((d, e), dtype=np.float32)
    for k in range(e):
        acc = None
        for j in range(e):
            if j != k and idx[j] == idx[k]:
                acc = src_np[:, j].copy() if acc is None else acc * src_np[:, j]
        ref[:, k] = dy[:, idx[k]].numpy() * (np.ones(d, dtype=np.float32) if acc is None else acc)
    assert np.array_equal(got.numpy(), ref)
    assert np.isfinite(got.numpy()).all()


def test_csr_and_transpose_layouts():
    rng = np.random.default_rng(2)
    n, e = 11, 60
    s, t = rng.integers(0, n, e), rng.integers(0, n, e)
    rowptr, ss, tt, perm = orc.csr_by_dst(s, t, n)
    assert np.all(np.diff(tt) >= 0) and np.array_equal(tt, t[perm]) and np.array_equal(ss, s[perm])
    for j in range(n):
        seg = perm[rowptr[j]:rowptr[j + 1]]
        assert np.array_equal(seg, np.nonzero(t == j)[0])  # ascending original position inside a row
    tptr, tpos = orc.csc_of_csr(ss, n)
    for i in range(n):
        seg = tpos[tptr[i]:tptr[i + 1]]
        assert np.array_equal(seg, np.nonzero(ss == i)[0])


def test_merged_adjacency_matches_dense():
    rng = np.random.default_rng(3)
    n, e = 7, 40
    s, t = rng.integers(0, n, e), rng.integers(0, n, e)
    w = rng.standard_normal(e)
    colptr, rowval, slot = orc.merged_adjacency(s, t, n)
    A = np.zeros((n, n))
    np.add.at(A, (s, t), w)
    val = np.zeros(len(rowval))
    np.add.at(val, slot, w)
    B = np.zeros((n, n))
    for c in range(n):
        rows = rowval[colptr[c]:colptr[c + 1]]
        assert np.all(np.diff(rows) > 0)  # ascending source, duplicates merged
        B[rows, c] = val[colptr[c]:colptr[c + 1]]
    assert np.allclose(A, B)


def test_greedy_units_cover_and_bound():
    rng = np.random.default_rng(4)
    deg = rng.integers(0, 40, 200)
    deg[17] = 500  # a row larger than any tile
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    for te in (32, 64, 128):
        u = orc.greedy_units(rowptr, te)
        assert u[0] == 0 and u[-1] == 200 and np.all(np.diff(u) >= 1) and np.all(np.diff(u) <= te)
        for a, b in zip(u[:-1], u[1:]):
            assert rowptr[b] - rowptr[a] <= te or b - a == 1


def test_reference_test_shapes():
    # test/runtests.jl:11-151 shape assertions on the 3-node toy graph, through the oracle
    rng = np.random.default_rng(0)
    s, t = np.array([0, 0, 1, 2]), np.array([1, 2, 0, 0])
    T = torch.float32
    g = orc.OGraph(s, t, 3)
    assert orc.gcn_conv(torch.randn(3, 3), {"weight": torch.randn(5, 3), "bias": torch.zeros(5, 1)}, g, 3, 5).shape == (5, 3)
    gh = orc.OGraph(s, t, 3, ndata={"x": torch.rand(3, 3)})
    phi, gam = [(11, 5, "identity", True)], [(9, 7, "identity", True)]
    u = torch.randn(4, 3)
    assert orc.explicit_edge_conv(u, orc.init_mlp(rng, phi), gh, phi).shape == (5, 3)
    ps = {"ϕ": orc.init_mlp(rng, phi), "γ": orc.init_mlp(rng, gam)}
    assert orc.vmh_conv(u, ps, gh, phi, gam).shape == (7, 3)
    # MPPDE: Float64 side data promotes the result (runtests.jl:58-61)
    gm = orc.OGraph(s, t, 3, ndata={"u": torch.rand(2, 3, dtype=torch.float64), "x": torch.rand(3, 3, dtype=torch.float64)},
                    gdata={"θ": torch.rand(4, dtype=torch.float64)})
    phi, psi = [(19, 5, "identity", True)], [(14, 7, "identity", True)]
    ps = {"ϕ": orc.init_mlp(rng, phi), "ψ": orc.init_mlp(rng, psi)}
    y = orc.mppde_conv(torch.randn(5, 3), ps, gm, phi, psi)
    assert y.shape == (7, 3) and y.dtype == torch.float64
    gb = orc.batch([gm, gm])
    assert orc.mppde_conv(torch.randn(5, 6), ps, gb, phi, psi).shape == (7, 6)
    # GNO on a graph whose last node is isolated: output still has num_nodes columns (runtests.jl:124-136)
    s2, t2 = rng.integers(0, 9, 6), rng.integers(0, 9, 6)
    gg = orc.OGraph(s2, t2, 10, ndata={"a": torch.rand(2, 10), "x": torch.rand(3, 10)})
    phi = [(10, 35, "identity", True)]
    ps = {"linear": {"weight": torch.randn(7, 5), "bias": torch.zeros(7, 1)}, "ϕ": orc.init_mlp(rng, phi)}
    assert orc.gno_conv(torch.randn(5, 10), ps, gg, 5, 7, phi).shape == (7, 10)
