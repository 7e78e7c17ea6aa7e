"""Synthetic graphs and layer configurations of BASELINE.json's `configs` (definitions: SURVEY.md section 8d).

Product-side only (numpy + the Lux mirror; nothing from `oracle/`).  Every workload returns the layer, its
parameters/state on the requested device, the node state `x` (Julia shape `(d, N)`) and the algorithmic work per
layer call, so bench.py, the parity tests and `__graft_entry__.smoke()` all exercise exactly the same objects.

    C1  ExplicitEdgeConv, 32x32 grid-4, phi 4=>16=>16=>1 (tanh), aggr mean            (launch-latency bound)
    C2  MPPDEConv, 64 x 256-node paths, phi 260=>128=>128, psi 258=>128=>128 (swish)   (compute bound)
    C3  VMHConv, 256x256 grid-8, phi 6=>64=>64=>64=>64, gamma 66=>64=>64=>64=>2 (tanh) (compute bound; the headline)
    C4  GNOConv 64=>64, random-geometric 1M nodes / ~16M edges, phi 6=>64=>64=>4096    (compute bound)
    C5  Chain(GCNConv(2=>64), GCNConv(64=>64), VMHConv(...)) on 512 x (64x64 grid-8)   (GCN aggregate: HBM bound)

`scale` shrinks a workload for tests (grid side / graph count / node count) without changing the model.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .graph import GNNGraph
from .layers import ExplicitEdgeConv, GCNConv, GNOConv, MPPDEConv, VMHConv
from .lux import Chain, Dense, setup


# ---------------------------------------------------------------------------------------------------------------
# graph generators (edge lists are emitted dst-major and then shuffled with the seeded generator, so that building
# the CSR layout is genuinely exercised -- SURVEY.md section 8d)
# ---------------------------------------------------------------------------------------------------------------

def grid_edges(nx: int, ny: int, neighbours: int = 8, rng: Optional[np.random.Generator] = None):
    idx = np.arange(nx * ny, dtype=np.int64).reshape(nx, ny)
    offs = [(-1, 0), (1, 0), (0, -1), (0, 1)]
    if neighbours == 8:
        offs += [(-1, -1), (-1, 1), (1, -1), (1, 1)]
    src, dst = [], []
    for di, dj in offs:
        i0, i1 = max(0, -di), min(nx, nx - di)
        j0, j1 = max(0, -dj), min(ny, ny - dj)
        dst.append(idx[i0:i1, j0:j1].ravel())
        src.append(idx[i0 + di:i1 + di, j0 + dj:j1 + dj].ravel())
    s, t = np.concatenate(src), np.concatenate(dst)
    o = np.argsort(t, kind="stable")
    s, t = s[o], t[o]
    if rng is not None:
        p = rng.permutation(len(s))
        s, t = s[p], t[p]
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    pos = np.stack([ii.ravel() / nx, jj.ravel() / nx]).astype(np.float32)  # (2, N)
    return s, t, pos


def path_edges(n_per: int, n_graphs: int, rng: Optional[np.random.Generator] = None):
    """`n_graphs` path graphs laid out contiguously, every graph with the same (shuffled) local edge order --
    MPPDEConv's `repeat(theta; inner=(1, E/G))` assumes graph-major edges (layers.jl:410)."""
    base = np.arange(n_per - 1, dtype=np.int64)
    s1, t1 = np.concatenate([base, base + 1]), np.concatenate([base + 1, base])
    if rng is not None:
        p = rng.permutation(len(s1))
        s1, t1 = s1[p], t1[p]
    offs = (np.arange(n_graphs, dtype=np.int64) * n_per)[:, None]
    return (s1[None, :] + offs).ravel(), (t1[None, :] + offs).ravel()


def radius_edges(n: int, mean_deg: float, rng: np.random.Generator):
    """Random-geometric graph: uniform points in [0,1]^2 sorted along a coarse strip/cell order (so a contiguous
    node range is a spatial strip: what the node partitioner wants), radius sqrt(mean_deg / (pi n))."""
    pos = rng.uniform(0.0, 1.0, size=(n, 2)).astype(np.float32)
    r = math.sqrt(mean_deg / (math.pi * n))
    nc = max(1, int(1.0 / r))
    cell = np.minimum((pos.astype(np.float64) * nc).astype(np.int64), nc - 1)
    cid = cell[:, 0] * nc + cell[:, 1]
    order = np.argsort(cid, kind="stable")
    pos, cell, cid = pos[order], cell[order], cid[order]
    start = np.searchsorted(cid, np.arange(nc * nc + 1))
    pd = pos.astype(np.float64)
    # candidate pairs cell-by-neighbour-cell, vectorised over all points of a cell row
    src, dst = [], []
    counts = np.diff(start)
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            cx, cy = cell[:, 0] + dx, cell[:, 1] + dy
            ok = (cx >= 0) & (cx < nc) & (cy >= 0) & (cy < nc)
            a = np.nonzero(ok)[0]
            nb = cx[a] * nc + cy[a]
            cnt = counts[nb]
            tot = int(cnt.sum())
            if tot == 0:
                continue
            rep = np.repeat(a, cnt)
            first = np.repeat(start[nb], cnt)
            within = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt)
            b = first + within
            d2 = ((pd[rep] - pd[b]) ** 2).sum(-1)
            keep = (d2 <= r * r) & (rep != b)
            dst.append(rep[keep])
            src.append(b[keep])
    s, t = np.concatenate(src).astype(np.int64), np.concatenate(dst).astype(np.int64)
    p = rng.permutation(len(s))
    return s[p], t[p], pos.T.copy()


# ---------------------------------------------------------------------------------------------------------------

@dataclass
class Workload:
    name: str
    layer: object
    ps: object
    st: object
    x: torch.Tensor                 # (d, N) Julia-shaped node state on the device
    graph: GNNGraph
    n_nodes: int
    n_edges: int
    flops_fwd: float                # E*F_edge + N*F_node
    bytes_fwd: float                # compulsory bytes of one fused forward
    bytes_fwdbwd: float
    rhs_per_step: int = 1
    notes: Dict[str, object] = field(default_factory=dict)


def _mlp_flops(spec: List[Tuple[int, int]]) -> int:
    return sum(2 * i * o for i, o in spec)


def _chain(dims: List[int], act: str, last_act: str = "identity") -> Chain:
    layers = []
    for i in range(len(dims) - 1):
        a = act if i < len(dims) - 2 else last_act
        layers.append(Dense(dims[i], dims[i + 1], a))
    return Chain(*layers)


def _x(rng, d, n, device):
    return torch.from_numpy(rng.uniform(-1.0, 1.0, size=(n, d)).astype(np.float32)).to(device).T


def _bytes(N, E, d_x, d_static, d_edata, d_out, n_params):
    fwd = 4 * (N * (d_x + d_static) + E * d_edata + N * d_out) + 4 * E + 4 * (N + 1) + 4 * n_params
    fwdbwd = fwd + 4 * (N * d_out + N * d_x) + 4 * E + 4 * (N + 1) + 4 * n_params
    return float(fwd), float(fwdbwd)


def c1_edgeconv(device="cuda", side: int = 32, seed: int = 0, hidden: int = 16) -> Workload:
    rng = np.random.default_rng(seed)
    s, t, pos = grid_edges(side, side, 4, rng)
    g = GNNGraph(torch.from_numpy(s), torch.from_numpy(t), num_nodes=side * side,
                 ndata={"x": torch.from_numpy(pos)}).to(device)
    layer = ExplicitEdgeConv(_chain([4, hidden, hidden, 1], "tanh"), initialgraph=g, aggr="mean")
    ps, st = setup(rng, layer, device)
    N, E = g.num_nodes, g.num_edges
    fe = _mlp_flops([(4, hidden), (hidden, hidden), (hidden, 1)])
    bf, bb = _bytes(N, E, 1, 2, 0, 1, layer.parameterlength())
    return Workload("C1 ExplicitEdgeConv %dx%d grid-4 h16" % (side, side), layer, ps, st, _x(rng, 1, N, device), g, N, E,
                    float(E * fe), bf, bb, rhs_per_step=6, notes={"flops_edge_fwd": float(E * fe), "flops_node_fwd": 0.0})


def c2_mppde(device="cuda", n_per: int = 256, n_graphs: int = 64, hidden: int = 128, seed: int = 0) -> Workload:
    rng = np.random.default_rng(seed)
    s, t = path_edges(n_per, n_graphs, rng)
    N = n_per * n_graphs
    u = rng.uniform(-1, 1, size=(1, N)).astype(np.float32)
    xs = np.tile(np.linspace(0.0, 1.0, n_per, dtype=np.float32), n_graphs)[None, :]
    theta = rng.uniform(-1, 1, size=(2, n_graphs)).astype(np.float32)
    g = GNNGraph(torch.from_numpy(s), torch.from_numpy(t), num_nodes=N, num_graphs=n_graphs,
                 ndata={"u": torch.from_numpy(u), "x": torch.from_numpy(xs)},
                 gdata={"θ": torch.from_numpy(theta)}).to(device)
    din_phi, din_psi = 2 * hidden + 2 + 2, 2 * hidden + 2
    layer = MPPDEConv(_chain([din_phi, hidden, hidden], "swish", "swish"),
                      _chain([din_psi, hidden, hidden], "swish", "swish"), initialgraph=g, aggr="mean")
    ps, st = setup(rng, layer, device)
    E = g.num_edges
    fe = _mlp_flops([(din_phi, hidden), (hidden, hidden)])
    fn = _mlp_flops([(din_psi, hidden), (hidden, hidden)])
    bf, bb = _bytes(N, E, hidden, 2, 0, hidden, layer.parameterlength())
    return Workload("C2 MPPDEConv %dx%d paths h%d" % (n_graphs, n_per, hidden), layer, ps, st, _x(rng, hidden, N, device),
                    g, N, E, float(E * fe + N * fn), bf, bb,
                    notes={"flops_edge_fwd": float(E * fe), "flops_node_fwd": float(N * fn)})


def c3_vmh(device="cuda", side: int = 256, hidden: int = 64, seed: int = 0) -> Workload:
    rng = np.random.default_rng(seed)
    s, t, pos = grid_edges(side, side, 8, rng)
    N = side * side
    g = GNNGraph(torch.from_numpy(s), torch.from_numpy(t), num_nodes=N, ndata={"x": torch.from_numpy(pos)}).to(device)
    h = hidden
    layer = VMHConv(_chain([6, h, h, h, h], "tanh"), _chain([h + 2, h, h, h, 2], "tanh"), initialgraph=g, aggr="mean")
    ps, st = setup(rng, layer, device)
    E = g.num_edges
    fe = _mlp_flops([(6, h), (h, h), (h, h), (h, h)])
    fn = _mlp_flops([(h + 2, h), (h, h), (h, h), (h, 2)])
    bf, bb = _bytes(N, E, 2, 2, 0, 2, layer.parameterlength())
    return Workload("C3 VMHConv %dx%d grid-8 h%d" % (side, side, h), layer, ps, st, _x(rng, 2, N, device), g, N, E,
                    float(E * fe + N * fn), bf, bb, rhs_per_step=4,
                    notes={"flops_edge_fwd": float(E * fe), "flops_node_fwd": float(N * fn)})


def c4_gno(device="cuda", n_nodes: int = 1_000_000, mean_deg: float = 16.0, chs: int = 64, hidden: int = 64,
           seed: int = 0) -> Workload:
    rng = np.random.default_rng(seed)
    s, t, pos = radius_edges(n_nodes, mean_deg, rng)
    a = rng.uniform(-1, 1, size=(1, n_nodes)).astype(np.float32)
    g = GNNGraph(torch.from_numpy(s), torch.from_numpy(t), num_nodes=n_nodes,
                 ndata={"a": torch.from_numpy(a), "x": torch.from_numpy(pos)}).to(device)
    phi = Chain(Dense(6, hidden, "relu"), Dense(hidden, hidden, "relu"), Dense(hidden, chs * chs))
    layer = GNOConv((chs, chs), phi, "relu", initialgraph=g, aggr="mean")
    ps, st = setup(rng, layer, device)
    N, E = n_nodes, g.num_edges
    fe = _mlp_flops([(6, hidden), (hidden, hidden), (hidden, chs * chs)]) + 2 * chs * chs
    fn = 2 * chs * chs
    bf, bb = _bytes(N, E, chs, 3, 0, chs, layer.parameterlength())
    # flops the factored evaluation (csrc/ngpde_gno.cuh) actually issues: hidden layers + outer product per edge, the
    # contraction with the affine last layer once per NODE (R = (hidden + 1) * chs rows of B = [W3; b3])
    R = (hidden + 1) * chs
    f_hidden = _mlp_flops([(6, hidden), (hidden, hidden)])
    ex_fwd = E * (f_hidden + 2 * R) + N * 2 * R * chs
    ex_bwd = E * (3 * f_hidden + 3 * 2 * R) + N * 2 * (2 * R * chs)
    return Workload("C4 GNOConv %d=>%d radius graph N=%d" % (chs, chs, N), layer, ps, st, _x(rng, chs, N, device), g, N, E,
                    float(E * fe + N * fn), bf, bb,
                    notes={"flops_edge_fwd": float(E * fe), "flops_node_fwd": float(N * fn),
                           "flops_executed_fwd_edge": float(ex_fwd), "flops_executed_bwd_edge": float(ex_bwd)})


def c5_gcn_vmh(device="cuda", n_graphs: int = 512, side: int = 64, hidden: int = 64, seed: int = 0) -> Workload:
    """ODE right-hand side (2 -> 2) of the ensemble config: GCNConv(2=>h,tanh), GCNConv(h=>h,tanh), VMHConv."""
    rng = np.random.default_rng(seed)
    s1, t1, pos1 = grid_edges(side, side, 8, rng)
    n1 = side * side
    offs = (np.arange(n_graphs, dtype=np.int64) * n1)[:, None]
    s, t = (s1[None, :] + offs).ravel(), (t1[None, :] + offs).ravel()
    pos = np.tile(pos1, (1, n_graphs))
    N = n1 * n_graphs
    g = GNNGraph(torch.from_numpy(s), torch.from_numpy(t), num_nodes=N, num_graphs=n_graphs,
                 ndata={"x": torch.from_numpy(pos)}).to(device)
    h = hidden
    vmh = VMHConv(Chain(Dense(2 * h + 2, h, "tanh"), Dense(h, h)), Chain(Dense(2 * h, h, "tanh"), Dense(h, 2)),
                  initialgraph=g, aggr="mean")
    layer = Chain(GCNConv((2, h), "tanh", initialgraph=g), GCNConv((h, h), "tanh", initialgraph=g), vmh)
    ps, st = setup(rng, layer, device)
    E = g.num_edges
    f_gcn = (E + N) * 2 * 2 + N * 2 * 2 * h + (E + N) * 2 * h + N * 2 * h * h
    fe = _mlp_flops([(2 * h + 2, h), (h, h)])
    fn = _mlp_flops([(2 * h, h), (h, 2)])
    bf, bb = _bytes(N, E, h, 2, 0, 2, layer.parameterlength())
    return Workload("C5 GCNConv+VMHConv %d x %dx%d grid-8 h%d" % (n_graphs, side, side, h), layer, ps, st,
                    _x(rng, 2, N, device), g, N, E, float(f_gcn + E * fe + N * fn), bf, bb, rhs_per_step=4,
                    notes={"flops_edge_fwd": float(E * fe), "flops_node_fwd": float(N * fn)})


WORKLOADS = {"c1": c1_edgeconv, "c2": c2_mppde, "c3": c3_vmh, "c4": c4_gno, "c5": c5_gcn_vmh}
