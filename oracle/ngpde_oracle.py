"""CPU oracle for the NeuralGraphPDE.jl message-passing hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.  The product package
(`neuralgraphpde.jl_b200/`) never imports anything from `oracle/`.

What it is: a from-spec restatement, in torch-CPU (float32 by default, float64 for gradient checks),
of the *unfused* algorithm the reference executes:

    gather(x, t), gather(x, s)  ->  vcat in the reference's row order  ->  Lux Dense chain
    ->  NNlib.scatter(aggr, m, t) as a sequential loop in stored edge order  ->  node update.

Parity status: the arithmetic of the hot path lives in un-vendored third-party Julia packages
(GraphNeuralNetworks.jl 0.4-0.6, NNlib 0.8, Lux 0.4; `/root/reference/Project.toml:20-32`), Julia is not
installed here, so the reference itself cannot be run.  The oracle is pinned against the only numerical
known-answer test the reference holds for this path -- SpectralConv, `/root/reference/test/runtests.jl:153-162`
and the doctest `/root/reference/src/layers.jl:581-631` (direction s->t, `+` aggregation) -- and against
the structural assertions of `test/runtests.jl:16-151`.  For the numerical results of the five layers and
for every gradient: **parity unpinned** (no golden vectors exist in the reference).

Array convention: Julia shapes.  A feature matrix is `(D, N)` (features x items); a Dense weight is
`(out, in)`, a bias `(out, 1)`.  Indices are 0-based here (the reference is 1-based).

Each function cites the reference file:line it restates.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# Graph container  ([DEP] GNNGraph COO semantics)
# --------------------------------------------------------------------------------------


@dataclass
class OGraph:
    """COO graph: edges stored in the order given, never sorted ([DEP] GNNGraph)."""

    s: np.ndarray  # (E,) int64, 0-based source of each edge
    t: np.ndarray  # (E,) int64, 0-based target of each edge
    num_nodes: int
    num_graphs: int = 1
    ndata: Dict[str, Tensor] = field(default_factory=dict)  # name -> (D, N)
    edata: Dict[str, Tensor] = field(default_factory=dict)  # name -> (D, E)
    gdata: Dict[str, Tensor] = field(default_factory=dict)  # name -> (D, G)
    w: Optional[Tensor] = None  # optional stored edge weights (E,)

    @property
    def num_edges(self) -> int:
        return int(self.s.shape[0])


def add_self_loops(g: OGraph) -> OGraph:
    """[DEP] GNNGraphs.add_self_loops: appends (i, i), i = 0..n-1, at the END of the COO lists
    (relied on at /root/reference/src/layers.jl:211-218)."""
    n = g.num_nodes
    loops = np.arange(n, dtype=np.int64)
    w = None
    if g.w is not None:
        w = torch.cat([g.w, torch.ones(n, dtype=g.w.dtype)])
    return OGraph(np.concatenate([g.s, loops]), np.concatenate([g.t, loops]), n, g.num_graphs,
                  g.ndata, g.edata, g.gdata, w)


def batch(gs: Sequence[OGraph]) -> OGraph:
    """[DEP] MLUtils.batch on GNNGraphs: block-diagonal concatenation, gdata concatenated along the
    last dim (/root/reference/test/runtests.jl:92)."""
    s, t, off = [], [], 0
    for g in gs:
        s.append(g.s + off)
        t.append(g.t + off)
        off += g.num_nodes
    cat = lambda attr: {k: torch.cat([getattr(g, attr)[k] for g in gs], dim=1) for k in getattr(gs[0], attr)}
    gd = {}
    for k in gs[0].gdata:
        parts = [g.gdata[k] if g.gdata[k].dim() == 2 else g.gdata[k].reshape(-1, 1) for g in gs]
        gd[k] = torch.cat(parts, dim=1)
    return OGraph(np.concatenate(s), np.concatenate(t), off, sum(g.num_graphs for g in gs),
                  cat("ndata"), cat("edata"), gd, None)


# --------------------------------------------------------------------------------------
# Index work: the layouts the CUDA library must reproduce bit-exactly
# --------------------------------------------------------------------------------------


def csr_by_dst(s: np.ndarray, t: np.ndarray, n: int):
    """Stable dst-sorted CSR.  perm[k] = original position of the k-th edge in CSR order, so within one
    destination row edges keep ascending original position == the order NNlib.scatter's sequential CPU loop
    visits them."""
    perm = np.argsort(t, kind="stable").astype(np.int64)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, t + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, s[perm].astype(np.int64), t[perm].astype(np.int64), perm


def csc_of_csr(src_sorted: np.ndarray, n: int):
    """Transpose layout used by the backward: CSR positions grouped by source, stable."""
    tperm = np.argsort(src_sorted, kind="stable").astype(np.int64)
    tptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(tptr, src_sorted + 1, 1)
    return np.cumsum(tptr), tperm


def in_degree(t: np.ndarray, n: int) -> np.ndarray:
    return np.bincount(t, minlength=n).astype(np.int64)


def merged_adjacency(s: np.ndarray, t: np.ndarray, n: int):
    """[DEP] `sparse(s, t, w, n, n)` as used by GNN.jl's adjacency_matrix: entries sorted by (column=t,
    row=s), duplicate (s, t) pairs merged.  Returns colptr, rowval and `slot[k]` = merged-entry index of
    original edge k (so weights can be summed in original order)."""
    key = t.astype(np.int64) * n + s.astype(np.int64)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    first = np.ones(len(ks), dtype=bool)
    first[1:] = ks[1:] != ks[:-1]
    slot_sorted = np.cumsum(first) - 1
    slot = np.empty(len(ks), dtype=np.int64)
    slot[order] = slot_sorted
    ukey = ks[first]
    rowval = (ukey % n).astype(np.int64)
    col = (ukey // n).astype(np.int64)
    colptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(colptr, col + 1, 1)
    return np.cumsum(colptr), rowval, slot


def greedy_units(rowptr: np.ndarray, te: int) -> np.ndarray:
    """Work units of the CUDA edge kernels: maximal runs of consecutive destination rows holding at most
    `te` edges and at most `te` rows; a row with more than `te` edges is a unit of its own."""
    n = len(rowptr) - 1
    bounds = [0]
    i = 0
    while i < n:
        j = i + 1
        while j < n and j - i < te and rowptr[j + 1] - rowptr[i] <= te:
            j += 1
        bounds.append(j)
        i = j
    return np.asarray(bounds, dtype=np.int64)


# --------------------------------------------------------------------------------------
# NNlib.gather / NNlib.scatter  ([DEP])
# --------------------------------------------------------------------------------------


def gather(x: Tensor, idx: np.ndarray) -> Tensor:
    """[DEP] NNlib.gather: dst[:, k] = src[:, idx[k]]."""
    return x.index_select(1, torch.from_numpy(np.ascontiguousarray(idx)))


def _rank_within_group(idx: np.ndarray, n: int) -> np.ndarray:
    """rank[k] = number of earlier positions k' < k with idx[k'] == idx[k]."""
    order = np.argsort(idx, kind="stable")
    counts = np.bincount(idx, minlength=n)
    starts = np.cumsum(counts) - counts
    rank = np.empty(len(idx), dtype=np.int64)
    rank[order] = np.arange(len(idx)) - starts[idx[order]]
    return rank


_IDENT = {"+": 0.0, "mean": 0.0, "max": -math.inf, "min": math.inf, "*": 1.0}


def _ordered_reduce(op: str, src: Tensor, idx: np.ndarray, n: int) -> Tensor:
    """Sequential loop `dst[:, idx[k]] = op(dst[:, idx[k]], src[:, k])` for k ascending, vectorised by
    rounds: round r applies the r-th edge of every destination, so the per-destination order of the
    floating-point operations is exactly the sequential one."""
    dst = torch.full((src.shape[0], n), _IDENT[op], dtype=src.dtype)
    if len(idx) == 0:
        return dst
    rank = _rank_within_group(idx, n)
    for r in range(int(rank.max()) + 1):
        sel = np.nonzero(rank == r)[0]
        cols = torch.from_numpy(idx[sel])
        cur = dst[:, cols]
        val = src[:, torch.from_numpy(sel)]
        if op in ("+", "mean"):
            new = cur + val
        elif op == "max":
            new = torch.maximum(cur, val)
        elif op == "min":
            new = torch.minimum(cur, val)
        elif op == "*":
            new = cur * val
        else:
            raise ValueError(op)
        dst[:, cols] = new
    return dst


def _prod_of_others(src: Tensor, idx: np.ndarray, n: int) -> Tensor:
    """po[:, k] = product of src[:, j] over the other positions j != k with idx[j] == idx[k], j ascending (left fold from
    the first factor; 1 when there is none).  Vectorised by rounds like _ordered_reduce: round r multiplies in the r-th
    member of every group for all members except that one."""
    po = torch.ones_like(src)
    if len(idx) == 0:
        return po
    rank = _rank_within_group(idx, n)
    order = np.argsort(idx, kind="stable")
    counts = np.bincount(idx, minlength=n)
    starts = np.cumsum(counts) - counts
    for r in range(int(rank.max()) + 1):
        has = counts[idx] > r                      # positions whose group has an r-th member
        sel = np.nonzero(has & (rank != r))[0]
        if len(sel) == 0:
            continue
        member = order[starts[idx[sel]] + r]       # position of the r-th member of each selected position's group
        sel_t = torch.from_numpy(sel)
        po[:, sel_t] = po[:, sel_t] * src[:, torch.from_numpy(member)]
    return po


class _Scatter(torch.autograd.Function):
    """[DEP] NNlib.scatter(op, src, idx; dstsize=(D, n)) and its ChainRules pullback w.r.t. src."""

    @staticmethod
    def forward(ctx, src: Tensor, idx_t: Tensor, n: int, op: str):
        idx = idx_t.numpy()
        s = src.detach()
        if op == "mean":
            tot = _ordered_reduce("+", s, idx, n)
            cnt = torch.from_numpy(np.bincount(idx, minlength=n).astype(np.float64)).to(s.dtype)
            out = torch.where(cnt > 0, tot / cnt, torch.zeros_like(tot))  # safe_div
            ctx.save_for_backward(idx_t, cnt)
        else:
            out = _ordered_reduce(op, s, idx, n)
            ctx.save_for_backward(idx_t, s, out)
        ctx.op = op
        return out

    @staticmethod
    def backward(ctx, gout: Tensor):
        op = ctx.op
        if op == "mean":
            idx_t, cnt = ctx.saved_tensors
            g = gout.index_select(1, idx_t) / cnt.index_select(0, idx_t)
        elif op == "+":
            idx_t, _, _ = ctx.saved_tensors
            g = gout.index_select(1, idx_t)
        elif op in ("max", "min"):
            idx_t, s, out = ctx.saved_tensors
            g = (s == out.index_select(1, idx_t)).to(gout.dtype) * gout.index_select(1, idx_t)
        else:  # "*": NNlib's ∇scatter_src for * -- gather(Δ, idx)[:, k] * prod(j -> src[:, j], others of idx[k]), the
            # others taken in ascending edge order (a left fold); never a division by src (zeros are exact)
            idx_t, s, out = ctx.saved_tensors
            g = gout.index_select(1, idx_t) * _prod_of_others(s, idx_t.numpy(), out.shape[1])
        return g, None, None, None


def scatter(op: str, src: Tensor, idx: np.ndarray, n: int) -> Tensor:
    return _Scatter.apply(src, torch.from_numpy(np.ascontiguousarray(idx)), n, op)


# --------------------------------------------------------------------------------------
# GraphNeuralNetworks.propagate  ([DEP], call sites /root/reference/src/layers.jl:111,326,416,534,656)
# --------------------------------------------------------------------------------------

Feat = Union[None, Tensor, Dict[str, Tensor]]


def _gather_feat(x: Feat, idx: np.ndarray) -> Feat:
    if x is None:
        return None
    if isinstance(x, dict):
        return {k: gather(v, idx) for k, v in x.items()}
    return gather(x, idx)


def apply_edges(f: Callable, g: OGraph, xi: Feat = None, xj: Feat = None, e: Feat = None):
    """xi <- features of the TARGET of each edge, xj <- features of the SOURCE."""
    return f(_gather_feat(xi, g.t), _gather_feat(xj, g.s), e)


def aggregate_neighbors(g: OGraph, aggr: str, m: Tensor) -> Tensor:
    return scatter(aggr, m, g.t, g.num_nodes)


def propagate(f: Callable, g: OGraph, aggr: str, xi: Feat = None, xj: Feat = None, e: Feat = None) -> Tensor:
    return aggregate_neighbors(g, aggr, apply_edges(f, g, xi, xj, e))


# --------------------------------------------------------------------------------------
# Lux Dense / Chain  ([DEP] Lux 0.4)
# --------------------------------------------------------------------------------------

_SQRT_2_OVER_PI = math.sqrt(2.0 / math.pi)

ACTIVATIONS: Dict[str, Callable[[Tensor], Tensor]] = {
    "identity": lambda x: x,
    "relu": torch.relu,
    "tanh": torch.tanh,  # Lux maps tanh -> NNlib.tanh_fast (a few ulp from tanh); exact tanh here
    "sigmoid": torch.sigmoid,
    "swish": lambda x: x * torch.sigmoid(x),
    "gelu": lambda x: 0.5 * x * (1.0 + torch.tanh(_SQRT_2_OVER_PI * (x + 0.044715 * x * x * x))),  # NNlib 0.8 gelu
    "softplus": torch.nn.functional.softplus,
    "elu": torch.nn.functional.elu,
    "leakyrelu": lambda x: torch.nn.functional.leaky_relu(x, 0.01),
}

# An MLP spec is a list of (in, out, activation-name, has_bias); its parameters are either a dict
# {"weight": (out,in), "bias": (out,1)} for a bare Dense, or {"layer_1": {...}, "layer_2": {...}} for a Chain.
MlpSpec = List[Tuple[int, int, str, bool]]


def dense(x: Tensor, p: Dict[str, Tensor], act: str, has_bias: bool = True) -> Tensor:
    """Lux.Dense: act.(weight * x .+ bias)."""
    W = p["weight"]
    dt = torch.promote_types(W.dtype, x.dtype)  # Julia promotes Float32 weights with Float64 side data
    y = W.to(dt) @ x.to(dt)
    if has_bias:
        y = y + p["bias"].to(dt)
    return ACTIVATIONS[act](y)


def mlp(x: Tensor, ps: Dict, spec: MlpSpec) -> Tensor:
    """Lux.Chain of Dense layers (or a single Dense when ps is un-nested)."""
    if "weight" in ps:
        (_, _, act, hb), = spec
        return dense(x, ps, act, hb)
    for i, (_, _, act, hb) in enumerate(spec):
        x = dense(x, ps[f"layer_{i + 1}"], act, hb)
    return x


def glorot_uniform(rng: np.random.Generator, out_dim: int, in_dim: int, dtype=np.float32) -> np.ndarray:
    """Lux.glorot_uniform(rng, out, in): U(-a, a), a = sqrt(24 / (in + out)) * 0.5 = sqrt(6/(in+out))."""
    a = math.sqrt(6.0 / (in_dim + out_dim))
    return rng.uniform(-a, a, size=(out_dim, in_dim)).astype(dtype)


def init_mlp(rng: np.random.Generator, spec: MlpSpec, dtype=torch.float32, chain: Optional[bool] = None) -> Dict:
    """Glorot-uniform weights, zero biases (Lux defaults; /root/reference/src/layers.jl:178-179,495)."""
    layers = {}
    for i, (din, dout, _, hb) in enumerate(spec):
        d = {"weight": torch.from_numpy(glorot_uniform(rng, dout, din)).to(dtype)}
        if hb:
            d["bias"] = torch.zeros(dout, 1, dtype=dtype)
        layers[f"layer_{i + 1}"] = d
    if chain is None:
        chain = len(spec) > 1
    return layers if chain else layers["layer_1"]


# --------------------------------------------------------------------------------------
# The five layers
# --------------------------------------------------------------------------------------


def _vcat(parts: Sequence[Tensor]) -> Tensor:
    dt = parts[0].dtype
    for p in parts:  # Julia promotes mixed Float32/Float64 (test/runtests.jl:58-61)
        dt = torch.promote_types(dt, p.dtype)
    return torch.cat([p.to(dt) for p in parts], dim=0)


def _as_named(x: Union[Tensor, Dict[str, Tensor]]) -> Dict[str, Tensor]:
    return x if isinstance(x, dict) else {"preservedname": x}


def explicit_edge_conv(x, ps: Dict, g: OGraph, phi: MlpSpec, aggr: str = "mean") -> Tensor:
    """/root/reference/src/layers.jl:94-112.  m = phi([h_i; h_j; pos_j - pos_i]); y = aggr_i(m)."""
    xs = dict(_as_named(x))
    xs.update(g.ndata)  # merge(x, s): later keys override, order = x's keys then new ndata keys

    def message(xi, xj, e):
        posi, posj = xi["x"], xj["x"]
        hi = [v for k, v in xi.items() if k != "x"]
        hj = [v for k, v in xj.items() if k != "x"]
        return mlp(_vcat(hi + hj + [posj - posi]), ps, phi)

    return propagate(message, g, aggr, xi=xs, xj=xs)


def vmh_conv(x, ps: Dict, g: OGraph, phi: MlpSpec, gamma: MlpSpec, aggr: str = "mean") -> Tensor:
    """/root/reference/src/layers.jl:308-332.  m = phi([h_i; h_j - h_i; pos_j - pos_i]); y = gamma([x; aggr(m)])."""
    xn = _as_named(x)
    xs = dict(xn)
    xs.update(g.ndata)

    def message(xi, xj, e):
        posi, posj = xi["x"], xj["x"]
        hi = [v for k, v in xi.items() if k != "x"]
        hj = [v for k, v in xj.items() if k != "x"]
        return mlp(_vcat(hi + [b - a for a, b in zip(hi, hj)] + [posj - posi]), ps["ϕ"], phi)

    m = propagate(message, g, aggr, xi=xs, xj=xs)
    return mlp(_vcat(list(xn.values()) + [m]), ps["γ"], gamma)


def _repeat_inner(theta: Tensor, reps: int) -> Tensor:
    """Julia repeat(θ; inner=(1, reps))."""
    return theta.repeat_interleave(reps, dim=1)


def mppde_conv(x: Tensor, ps: Dict, g: OGraph, phi: MlpSpec, psi: MlpSpec, aggr: str = "mean") -> Tensor:
    """/root/reference/src/layers.jl:390-422.
    m = phi([h_i; h_j; s_i - s_j; e_ij; θ]); y = psi([x; aggr(m); θ]); θ has no gradient (:397, :418)."""
    E, N, G = g.num_edges, g.num_nodes, g.num_graphs
    thetas = [v if v.dim() == 2 else v.reshape(-1, 1) for v in g.gdata.values()]
    theta = _vcat(thetas).detach() if thetas else torch.zeros(0, G, dtype=x.dtype)
    nkeys = list(g.ndata.keys())

    def message(xi, xj, e):
        parts = [xi["preservedname"], xj["preservedname"]]
        if nkeys:
            di = _vcat([xi[k] for k in nkeys])
            dj = _vcat([xj[k] for k in nkeys])
            parts.append(di - dj)
        if e:
            parts.append(_vcat(list(e.values())))
        if theta.shape[0] > 0:
            parts.append(_repeat_inner(theta, E // G))
        return mlp(_vcat(parts), ps["ϕ"], phi)

    xs = {"preservedname": x}
    xs.update(g.ndata)
    m = propagate(message, g, aggr, xi=xs, xj=xs, e=g.edata)
    parts = [x, m]
    if theta.shape[0] > 0:
        parts.append(_repeat_inner(theta, N // G))
    return mlp(_vcat(parts), ps["ψ"], psi)


def gno_conv(x: Tensor, ps: Dict, g: OGraph, in_chs: int, out_chs: int, phi: MlpSpec, act: str = "identity",
             aggr: str = "mean", bias: bool = True) -> Tensor:
    """/root/reference/src/layers.jl:509-547.
    W_k = reshape(phi([s_i; s_j; e_k]), out, in) (column-major: flat index o + out*i);
    m_k = W_k h_j;  y = act(W_lin x + aggr(m) + b)."""
    E = g.num_edges
    nkeys = list(g.ndata.keys())

    def message(xi, xj, e):
        parts = []
        if nkeys:
            parts.append(_vcat([xi[k] for k in nkeys]))
            parts.append(_vcat([xj[k] for k in nkeys]))
        if e:
            parts.append(_vcat(list(e.values())))
        W = mlp(_vcat(parts), ps["ϕ"], phi)  # (in*out, E)
        hj = xj["h_"]
        # Julia reshape(W, :, in, E): W3[o, i, k] = W[o + out*i, k]
        W3 = W.reshape(in_chs, out_chs, E).permute(1, 0, 2)  # (out, in, E)
        return torch.einsum("oik,ik->ok", W3, hj.to(W3.dtype))

    xs = {"h_": x}
    xs.update(g.ndata)
    m = propagate(message, g, aggr, xi=xs, xj=xs, e=g.edata)
    y = ps["linear"]["weight"] @ x + m
    if bias:
        y = y + ps["linear"]["bias"]
    return ACTIVATIONS[act](y)


class _SpmmOrdered(torch.autograd.Function):
    """[DEP] `xj * A`, A = SparseMatrixCSC(s, t, w) with duplicates merged: for each column (destination)
    C[:, col] += X[:, rowval[k]] * nzval[k] for k ascending (ascending SOURCE index), multiply and add rounded
    separately."""

    @staticmethod
    def forward(ctx, x: Tensor, val: Tensor, colptr_t: Tensor, rowval_t: Tensor):
        colptr, rowval = colptr_t.numpy(), rowval_t.numpy()
        n = len(colptr) - 1
        col = np.repeat(np.arange(n), np.diff(colptr))
        out = torch.zeros(x.shape[0], n, dtype=x.dtype)
        if len(rowval):
            rank = np.arange(len(rowval)) - colptr[col]
            xd, vd = x.detach(), val.detach()
            for r in range(int(rank.max()) + 1):
                sel = np.nonzero(rank == r)[0]
                c = torch.from_numpy(col[sel])
                prod = xd[:, torch.from_numpy(rowval[sel])] * vd[torch.from_numpy(sel)]
                out[:, c] = out[:, c] + prod
        ctx.save_for_backward(x, val, torch.from_numpy(col), rowval_t)
        return out

    @staticmethod
    def backward(ctx, gout):
        x, val, col_t, rowval_t = ctx.saved_tensors
        gx = torch.zeros_like(x)
        gx.index_add_(1, rowval_t, gout.index_select(1, col_t) * val)
        gval = (gout.index_select(1, col_t) * x.index_select(1, rowval_t)).sum(0)
        return gx, gval, None, None


def gcn_conv(x: Tensor, ps: Dict, g: OGraph, in_chs: int, out_chs: int, act: str = "identity",
             add_loops: bool = True, use_edge_weight: bool = False, edge_weight: Optional[Tensor] = None,
             bias: bool = True) -> Tensor:
    """/root/reference/src/layers.jl:200-239 on the CPU sparse-matmul path of GNN.jl ([DEP])."""
    if edge_weight is not None:
        assert edge_weight.shape[0] == g.num_edges, \
            f"Wrong number of edge weights (expected {g.num_edges} but given {edge_weight.shape[0]})"  # :207
    if add_loops:
        g = add_self_loops(g)  # :211
        if edge_weight is not None:
            edge_weight = torch.cat([edge_weight, torch.ones(g.num_nodes, dtype=edge_weight.dtype)])  # :215
    n = g.num_nodes
    if out_chs < in_chs:
        x = ps["weight"] @ x  # :220-223
    # degree(g, T; dir=:in, edge_weight) (:224): the explicit vector when one was passed; with edge_weight === nothing
    # GNN.jl's `_get_edge_weight(g, nothing)` resolves to the graph's own stored weights (self-loops padded with ones by
    # add_self_loops), whatever use_edge_weight says; a graph without weights gives the plain in-degree
    if edge_weight is not None:
        d = scatter("+", edge_weight.reshape(1, -1).to(x.dtype), g.t, n).reshape(-1)
    elif g.w is not None:
        d = scatter("+", g.w.reshape(1, -1).to(x.dtype), g.t, n).reshape(-1)
    else:
        d = torch.from_numpy(in_degree(g.t, n).astype(np.float64)).to(x.dtype)
    c = 1.0 / torch.sqrt(d)  # :225
    x = x * c  # :226  (x .* c')
    colptr, rowval, slot = merged_adjacency(g.s, g.t, n)
    if edge_weight is not None:  # e_mul_xj  :228
        w = edge_weight.to(x.dtype)
    elif use_edge_weight and g.w is not None:  # w_mul_xj  :230
        w = g.w.to(x.dtype)
    else:  # copy_xj  :232  (w_mul_xj on a graph without weights is the same thing)
        w = torch.ones(g.num_edges, dtype=x.dtype)
    val = scatter("+", w.reshape(1, -1), slot, len(rowval)).reshape(-1)  # duplicates summed, original order
    x = _SpmmOrdered.apply(x, val, torch.from_numpy(colptr), torch.from_numpy(rowval))
    x = x * c  # :234
    if out_chs >= in_chs:
        x = ps["weight"] @ x  # :235-237
    if bias:
        x = x + ps["bias"]  # :238 (the reference reads ps.bias unconditionally; bias=false would error there)
    return ACTIVATIONS[act](x)


# --------------------------------------------------------------------------------------
# SpectralConv: the reference's only numerical known-answer test for propagate
# --------------------------------------------------------------------------------------


def spectral_graph(n: int, dtype=torch.float64) -> OGraph:
    """/root/reference/src/layers.jl:639-648: complete digraph, edata e = x[t] - x[s]."""
    ii, jj = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    mask = ii != jj
    s, t = ii[mask].astype(np.int64), jj[mask].astype(np.int64)  # Graphs.edges order: src-major
    xs = np.linspace(0.0, 2.0 * math.pi, n + 1)[1:]
    diff = torch.from_numpy(xs[t] - xs[s]).to(dtype).reshape(1, -1)
    return OGraph(s, t, n, 1, {}, {"e": diff}, {})


def spectral_conv(x: Tensor, g: OGraph, n: int) -> Tensor:
    """/root/reference/src/layers.jl:652-657."""

    def message(xi, xj, e):
        return torch.cos(e * n / 2) * (1.0 / torch.tan(e / 2)) / 2 * xj

    return propagate(message, g, "+", xj=x, e=g.edata["e"])


# --------------------------------------------------------------------------------------
# Synthetic graphs of SURVEY.md §8(d)
# --------------------------------------------------------------------------------------


def grid_graph(nx: int, ny: int, neighbours: int = 8, rng: Optional[np.random.Generator] = None):
    """nx x ny lattice, 4- or 8-neighbour, directed both ways, positions (i, j)/nx.  Edges are emitted
    dst-major and then shuffled with `rng` (if given) so layout building is exercised."""
    idx = np.arange(nx * ny).reshape(nx, ny)
    offs = [(-1, 0), (1, 0), (0, -1), (0, 1)]
    if neighbours == 8:
        offs += [(-1, -1), (-1, 1), (1, -1), (1, 1)]
    s_l, t_l = [], []
    for di, dj in offs:
        i0, i1 = max(0, -di), min(nx, nx - di)
        j0, j1 = max(0, -dj), min(ny, ny - dj)
        t_l.append(idx[i0:i1, j0:j1].ravel())
        s_l.append(idx[i0 + di:i1 + di, j0 + dj:j1 + dj].ravel())
    s, t = np.concatenate(s_l).astype(np.int64), np.concatenate(t_l).astype(np.int64)
    o = np.argsort(t, kind="stable")
    s, t = s[o], t[o]
    if rng is not None:
        p = rng.permutation(len(s))
        s, t = s[p], t[p]
    ii, jj = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    pos = np.stack([ii.ravel() / nx, jj.ravel() / nx]).astype(np.float32)  # (2, N)
    return s, t, pos


def path_graphs(n_per: int, n_graphs: int, rng: Optional[np.random.Generator] = None):
    """`n_graphs` path graphs of `n_per` nodes (+-1 neighbour), block-diagonal, graph-major edge order."""
    base = np.arange(n_per - 1)
    s1 = np.concatenate([base, base + 1])
    t1 = np.concatenate([base + 1, base])
    if rng is not None:
        p = rng.permutation(len(s1))
        s1, t1 = s1[p], t1[p]
    s = np.concatenate([s1 + k * n_per for k in range(n_graphs)]).astype(np.int64)
    t = np.concatenate([t1 + k * n_per for k in range(n_graphs)]).astype(np.int64)
    return s, t


def radius_graph(n: int, mean_deg: float, rng: np.random.Generator):
    """Random-geometric graph, uniform in [0,1]^2, radius sqrt(mean_deg/(pi n)), via a cell grid."""
    pos = rng.uniform(0.0, 1.0, size=(n, 2)).astype(np.float32)
    r = math.sqrt(mean_deg / (math.pi * n))
    nc = max(1, int(1.0 / r))
    cell = np.minimum((pos / (1.0 / nc)).astype(np.int64), nc - 1)
    cid = cell[:, 0] * nc + cell[:, 1]
    order = np.argsort(cid, kind="stable")
    start = np.searchsorted(cid[order], np.arange(nc * nc + 1))
    s_l, t_l = [], []
    pd = pos.astype(np.float64)
    for cx in range(nc):
        for cy in range(nc):
            a = order[start[cx * nc + cy]:start[cx * nc + cy + 1]]
            if len(a) == 0:
                continue
            nb = []
            for dx in (-1, 0, 1):
                for dy in (-1, 0, 1):
                    x2, y2 = cx + dx, cy + dy
                    if 0 <= x2 < nc and 0 <= y2 < nc:
                        nb.append(order[start[x2 * nc + y2]:start[x2 * nc + y2 + 1]])
            b = np.concatenate(nb)
            d2 = ((pd[a, None, :] - pd[None, b, :]) ** 2).sum(-1)
            ia, ib = np.nonzero((d2 <= r * r) & (a[:, None] != b[None, :]))
            t_l.append(a[ia])
            s_l.append(b[ib])
    s, t = np.concatenate(s_l).astype(np.int64), np.concatenate(t_l).astype(np.int64)
    p = rng.permutation(len(s))
    return s[p], t[p], pos.T.copy()


# --------------------------------------------------------------------------------------
# Fixed-step Runge-Kutta restatement ([DEP] DifferentialEquations.jl `RK4()` / `Tsit5()` with adaptive=false;
# call sites /root/reference/docs/src/tutorials/graph_node.md:59-66, VMH.md:87).  Plain torch-CPU axpys.
# --------------------------------------------------------------------------------------

RK4_TABLEAU = (((), (0.5,), (0.0, 0.5), (0.0, 0.0, 1.0)), (1 / 6, 1 / 3, 1 / 3, 1 / 6))
TSIT5_TABLEAU = (
    ((), (0.161,), (-0.008480655492356989, 0.335480655492357),
     (2.8971530571054935, -6.359448489975075, 4.3622954328695815),
     (5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525),
     (5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383)),
    (0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774),
)


def integrate_fixed(rhs: Callable[[Tensor], Tensor], u0: Tensor, t0: float, t1: float, dt: float,
                    method: str = "rk4") -> Tensor:
    A, b = RK4_TABLEAU if method == "rk4" else TSIT5_TABLEAU
    u = u0
    for _ in range(int(round((t1 - t0) / dt))):
        ks: List[Tensor] = []
        for row in A:
            ui = u
            for a, k in zip(row, ks):
                if a != 0.0:
                    ui = ui + (dt * a) * k
            ks.append(rhs(ui))
        for w, k in zip(b, ks):
            u = u + (dt * w) * k
    return u


# --------------------------------------------------------------------------------------
# the training step either side of the adjoint.  The reference reaches these through un-vendored
# packages (Optimisers.jl, Flux.Losses / its own one-liner); their published rules are restated here in float32 numpy with
# one rounding per operation, in the operation order of the Julia source.
# --------------------------------------------------------------------------------------


def adam_init(x: np.ndarray, beta=(0.9, 0.999)):
    """Optimisers.jl `init(o::Adam, x) = (zero(x), zero(x), o.beta)` (call site: graph_node.md:122-123)."""
    return np.zeros_like(x), np.zeros_like(x), (np.float32(beta[0]), np.float32(beta[1]))


def adam_step(x: np.ndarray, g: np.ndarray, state, eta=0.001, beta=(0.9, 0.999), eps=1e-8):
    """Optimisers.jl `apply!(o::Adam, state, x, dx)` followed by `x .- dx'`:
        mt = b1 mt + (1 - b1) dx;  vt = b2 vt + (1 - b2) dx^2;  dx' = mt / (1 - b1^t) / (sqrt(vt / (1 - b2^t)) + eps) * eta."""
    f = np.float32
    m, v, bt = state
    b1, b2, eta, eps = f(beta[0]), f(beta[1]), f(eta), f(eps)
    m = b1 * m + (f(1) - b1) * g
    v = b2 * v + (f(1) - b2) * (g * g)
    step = m / (f(1) - bt[0]) / (np.sqrt(v / (f(1) - bt[1])) + eps) * eta
    return (x - step).astype(np.float32), (m.astype(np.float32), v.astype(np.float32), (f(bt[0] * b1), f(bt[1] * b2)))


def rprop_init(x: np.ndarray, eta=1e-3):
    """Optimisers.jl `init(o::Rprop, x) = (zero(x), onevalue(o.eta, x))` (call site: VMH.md:97)."""
    return np.zeros_like(x), np.full_like(x, np.float32(eta))


def rprop_step(x: np.ndarray, dx: np.ndarray, state, ell=(0.5, 1.2), gamma=(1e-6, 50.0)):
    """Optimisers.jl `apply!(o::Rprop, state, x, dx)`: eta grows by ell[2] (capped at gamma[2]) while g*dx > 0, shrinks by
    ell[1] (floored at gamma[1]) when g*dx < 0, in which case the stored gradient is zeroed; x -= eta * sign(g)."""
    f = np.float32
    g, eta = state
    prod = g * dx
    eta = np.where(prod > 0, np.minimum(eta * f(ell[1]), f(gamma[1])),
                   np.where(prod < 0, np.maximum(eta * f(ell[0]), f(gamma[0])), eta)).astype(np.float32)
    g = np.where(prod < 0, f(0), dx).astype(np.float32)
    return (x - eta * np.sign(g)).astype(np.float32), (g, eta)


def mse(yhat: Tensor, y: Tensor) -> Tensor:
    """Flux.Losses.mse (VMH.md:105-109): mean(abs2, yhat - y)."""
    return ((yhat - y) ** 2).mean()


def logitcrossentropy(yhat: Tensor, y: Tensor) -> Tensor:
    """graph_node.md:100: `mean(-sum(y .* logsoftmax(yhat); dims=1))` on Julia-shaped (classes, items) arrays."""
    return (-(y * torch.log_softmax(yhat, dim=0)).sum(dim=0)).mean()
