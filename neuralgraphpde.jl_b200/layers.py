"""The five message-passing layers of NeuralGraphPDE.jl behind the unchanged Lux calling convention

    ps, st = setup(rng, layer);   y, st = layer(x, ps, st)        with st.graph :: GNNGraph

(/root/reference/src/layers.jl).  Constructors, keyword names, parameter/state trees and error behaviour follow the
reference; the bodies call the fused CUDA kernels of libngpde through ops.py.  `x`, `y`, weights and graph data use
Julia shapes `(features, items)`, stored column-major.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Sequence, Union

import torch

from . import _lib, ops
from .graph import GNNGraph, copy, from_rowmajor, rowmajor
from .lux import (NT, AbstractExplicitLayer, Chain, ComponentArray, Dense, _act_name, flat_params, glorot_normal,
                  glorot_uniform, julia_array, merge, mlp_spec, nfkc, zeros32)

Tensor = torch.Tensor

_EMPTYGRAPH = None


def initialgraph() -> GNNGraph:
    """Default `initialgraph`: the shared empty graph (layers.jl:14,21)."""
    global _EMPTYGRAPH
    if _EMPTYGRAPH is None:
        z = torch.zeros(0, dtype=torch.int64)
        _EMPTYGRAPH = GNNGraph(z, z.clone(), num_nodes=0)
    return _EMPTYGRAPH


def wrapgraph(g: Union[GNNGraph, Callable[[], GNNGraph]]) -> Callable[[], GNNGraph]:
    """wrapgraph(g::GNNGraph) = () -> copy(g); wrapgraph(f::Function) = f  (/root/reference/src/utils.jl:16-17)."""
    if isinstance(g, GNNGraph):
        return lambda: copy(g)
    return g


def _aggr_name(aggr) -> str:
    if callable(aggr):
        aggr = {torch.mean: "mean", torch.sum: "+", sum: "+", max: "max", min: "min", torch.max: "max",
                torch.min: "min", torch.prod: "*"}.get(aggr, getattr(aggr, "__name__", str(aggr)))
    if aggr not in _lib.AGGR:
        raise ValueError(f"unsupported aggregation {aggr!r}; supported: +, *, mean, max, min")
    return aggr


class AbstractGNNLayer(AbstractExplicitLayer):
    """layers.jl:5,23: st = (graph = l.initialgraph(),)."""

    def initialstates(self, rng) -> NT:
        return NT(graph=self.initialgraph())

    def statelength(self) -> int:
        return 1


class AbstractGNNContainerLayer(AbstractExplicitLayer):
    """layers.jl:12,26-34: one (empty) sub-state per sub-layer, then the graph.  Parameters follow the Lux container
    rule: a single sub-layer is returned un-nested (docs/src/devdoc.md:74-88)."""

    sublayers: Sequence[str] = ()

    def initialstates(self, rng) -> NT:
        st = NT((name, getattr(self, nfkc(name)).initialstates(rng)) for name in self.sublayers)
        dict.__setitem__(st, "graph", self.initialgraph())
        return st

    def initialparameters(self, rng, device="cpu") -> NT:
        if len(self.sublayers) == 1:
            return getattr(self, nfkc(self.sublayers[0])).initialparameters(rng, device)
        return NT((name, getattr(self, nfkc(name)).initialparameters(rng, device)) for name in self.sublayers)

    def parameterlength(self) -> int:
        return sum(getattr(self, nfkc(n)).parameterlength() for n in self.sublayers)

    def statelength(self) -> int:
        return sum(getattr(self, nfkc(n)).statelength() for n in self.sublayers) + 1


def _named(x) -> Dict[str, Tensor]:
    return dict(x) if isinstance(x, dict) else {"preservedname": x}


def _concat_rm(parts: Sequence[Tensor]) -> Tensor:
    rm = [rowmajor(p) for p in parts]
    return rm[0] if len(rm) == 1 else torch.cat(rm, dim=1)


def _conv_desc(family: str, aggr: str, dx: int, dhs: int, dpos: int, de: int, dtheta: int, phi, node,
               gno_in: int = 0, gno_out: int = 0) -> _lib.ConvDesc:
    d = _lib.ConvDesc()
    d.family, d.aggr = _lib.FAMILY[family], _lib.AGGR[aggr]
    d.dx, d.dhs, d.dpos, d.de, d.dtheta = dx, dhs, dpos, de, dtheta
    d.gno_in, d.gno_out = gno_in, gno_out
    d.phi = _lib.make_mlp(phi)
    d.node = _lib.make_mlp(node) if node else _lib.Mlp()
    return d


def _nparams(spec) -> int:
    return sum(i * o + (o if b else 0) for i, o, _, b in spec)


def _reject_collisions(xs: Dict[str, Tensor], g: GNNGraph, who: str):
    """`xs = merge(x, g.ndata)` (layers.jl:110,324): on a key present in both, Julia's merge keeps x's position but takes
    the ndata VALUE, i.e. the trainable input would be silently replaced by static data (and, in VMHConv, gamma would
    still see the input's value).  The fused kernels keep trainable and static columns apart, so this corner is refused
    loudly instead of being computed differently from the reference."""
    both = [k for k in xs if k in g.ndata]
    if both:
        raise _lib.NgpdeError(f"{who}: field(s) {both} appear both in the input NamedTuple and in st.graph.ndata; "
                              "merge(x, ndata) would replace the input by the static data (layers.jl:110,324) -- rename one")


def _check_nodes(x_rm: Tensor, g: GNNGraph):
    if x_rm.shape[0] != g.num_nodes:
        raise ValueError(f"DimensionMismatch: x has {x_rm.shape[0]} columns but the graph has {g.num_nodes} nodes")


class ExplicitEdgeConv(AbstractGNNContainerLayer):
    """h'_i = aggr_j phi([h_i; h_j; x_j - x_i])  (layers.jl:84-112).  `ps` are phi's parameters, un-nested."""

    sublayers = ("ϕ",)

    def __init__(self, ϕ: AbstractExplicitLayer, *, initialgraph=initialgraph, aggr="mean"):
        self.ϕ = ϕ
        self.initialgraph = wrapgraph(initialgraph)
        self.aggr = _aggr_name(aggr)

    def prepare(self, x, ps, st: NT):
        """Everything one C-ABI call needs: (x [N,dx], phi params, node params, handle, desc, snode, edata, theta, dm, dy)."""
        g: GNNGraph = st.graph
        xs = _named(x)
        if "x" not in g.ndata:
            raise KeyError("ExplicitEdgeConv needs the spatial coordinates in st.graph.ndata.x (layers.jl:98-105)")
        dev = next(iter(xs.values())).device
        _reject_collisions(xs, g, "ExplicitEdgeConv")
        x_rm = _concat_rm([v for k, v in xs.items() if k != "x"])
        _check_nodes(x_rm, g)
        hs_keys = [k for k in g.ndata if k != "x"]
        snode = g._packed("edgeconv_s", hs_keys + ["x"], g.ndata, dev)
        dhs = snode.shape[1] - g.ndata["x"].shape[0]
        phi = mlp_spec(self.ϕ)
        desc = _conv_desc("explicit_edge_conv", self.aggr, x_rm.shape[1], dhs, g.ndata["x"].shape[0], 0, 0, phi, None)
        return (x_rm, flat_params(ps, _nparams(phi)), None, g.handle(dev), desc, snode, None, None, phi[-1][1], phi[-1][1])

    def __call__(self, x, ps, st: NT):
        return from_rowmajor(ops.ConvFunction.apply(*self.prepare(x, ps, st))), st


class VMHConv(AbstractGNNContainerLayer):
    """m_i = aggr_j phi([h_i; h_j - h_i; x_j - x_i]);  h'_i = gamma([h_i; m_i])  (layers.jl:295-332)."""

    sublayers = ("ϕ", "γ")

    def __init__(self, ϕ: AbstractExplicitLayer, γ: AbstractExplicitLayer, *, initialgraph=initialgraph, aggr="mean"):
        self.ϕ, self.γ = ϕ, γ
        self.initialgraph = wrapgraph(initialgraph)
        self.aggr = _aggr_name(aggr)

    def prepare(self, x, ps, st: NT):
        g: GNNGraph = st.graph
        xs = _named(x)
        if "x" not in g.ndata:
            raise KeyError("VMHConv needs the spatial coordinates in st.graph.ndata.x (layers.jl:313-321)")
        dev = next(iter(xs.values())).device
        _reject_collisions(xs, g, "VMHConv")
        x_rm = _concat_rm([v for k, v in xs.items() if k != "x"])
        _check_nodes(x_rm, g)
        hs_keys = [k for k in g.ndata if k != "x"]
        snode = g._packed("vmh_s", hs_keys + ["x"], g.ndata, dev)
        dpos = g.ndata["x"].shape[0]
        phi, gamma = mlp_spec(self.ϕ), mlp_spec(self.γ)
        desc = _conv_desc("vmh_conv", self.aggr, x_rm.shape[1], snode.shape[1] - dpos, dpos, 0, 0, phi, gamma)
        return (x_rm, flat_params(ps.ϕ, _nparams(phi)), flat_params(ps.γ, _nparams(gamma)), g.handle(dev), desc, snode,
                None, None, phi[-1][1], gamma[-1][1])

    def __call__(self, x, ps, st: NT):
        return from_rowmajor(ops.ConvFunction.apply(*self.prepare(x, ps, st))), st


class MPPDEConv(AbstractGNNContainerLayer):
    """m_i = aggr_j phi([h_i; h_j; s_i - s_j; e_ij; θ]);  h'_i = psi([h_i; m_i; θ])  (layers.jl:377-422)."""

    sublayers = ("ϕ", "ψ")

    def __init__(self, ϕ: AbstractExplicitLayer, ψ: AbstractExplicitLayer, *, aggr="mean", initialgraph=initialgraph):
        self.ϕ, self.ψ = ϕ, ψ
        self.initialgraph = wrapgraph(initialgraph)
        self.aggr = _aggr_name(aggr)

    def prepare(self, x: Tensor, ps, st: NT):
        g: GNNGraph = st.graph
        dev = x.device
        x_rm = rowmajor(x)
        _check_nodes(x_rm, g)
        snode = g._packed("mppde_s", list(g.ndata), g.ndata, dev)
        edata = g._packed("mppde_e", list(g.edata), g.edata, dev)
        gd = {k: (v if v.dim() == 2 else v.reshape(-1, 1)) for k, v in g.gdata.items()}
        theta = g._packed("mppde_t", list(gd), gd, dev)
        if theta is not None and theta.shape[0] != g.num_graphs:
            raise ValueError(f"gdata has {theta.shape[0]} columns but the graph batch holds {g.num_graphs} graphs")
        phi, psi = mlp_spec(self.ϕ), mlp_spec(self.ψ)
        desc = _conv_desc("mppde_conv", self.aggr, x_rm.shape[1], 0 if snode is None else snode.shape[1], 0,
                          0 if edata is None else edata.shape[1], 0 if theta is None else theta.shape[1], phi, psi)
        return (x_rm, flat_params(ps.ϕ, _nparams(phi)), flat_params(ps.ψ, _nparams(psi)), g.handle(dev), desc, snode,
                edata, theta, phi[-1][1], psi[-1][1])

    def __call__(self, x: Tensor, ps, st: NT):
        return from_rowmajor(ops.ConvFunction.apply(*self.prepare(x, ps, st))), st


class GNOConv(AbstractGNNContainerLayer):
    """m_i = aggr_j reshape(phi([s_i; s_j; e_ij]), out, in) h_j;  h'_i = act(W h_i + m_i + b)  (layers.jl:485-547)."""

    sublayers = ("linear", "ϕ")

    def __init__(self, in_chs, out_chs=None, ϕ: AbstractExplicitLayer = None, activation="identity", *,
                 initialgraph=initialgraph, init_weight=glorot_uniform, init_bias=zeros32, aggr="mean",
                 bias: bool = True):
        if isinstance(in_chs, tuple):  # GNOConv((in, out), ϕ, act) stands for Julia's `in => out`
            if ϕ is not None:
                activation = ϕ
            in_chs, out_chs, ϕ = in_chs[0], in_chs[1], out_chs
        self.in_chs, self.out_chs = int(in_chs), int(out_chs)
        self.initialgraph = wrapgraph(initialgraph)
        self.aggr = _aggr_name(aggr)
        self.bias = bool(bias)
        self.linear = Dense(self.in_chs, self.out_chs, activation, bias=bias, init_weight=init_weight,
                            init_bias=init_bias)
        self.ϕ = ϕ

    def prepare(self, x: Tensor, ps, st: NT):
        g: GNNGraph = st.graph
        dev = x.device
        x_rm = rowmajor(x)
        _check_nodes(x_rm, g)
        snode = g._packed("gno_s", list(g.ndata), g.ndata, dev)
        edata = g._packed("gno_e", list(g.edata), g.edata, dev)
        phi, lin = mlp_spec(self.ϕ), self.linear.spec()
        desc = _conv_desc("gno_conv", self.aggr, x_rm.shape[1], 0 if snode is None else snode.shape[1], 0,
                          0 if edata is None else edata.shape[1], 0, phi, lin, self.in_chs, self.out_chs)
        return (x_rm, flat_params(ps.ϕ, _nparams(phi)), flat_params(ps.linear, _nparams(lin)), g.handle(dev), desc, snode,
                edata, None, self.out_chs, self.out_chs)

    def __call__(self, x: Tensor, ps, st: NT):
        return from_rowmajor(ops.ConvFunction.apply(*self.prepare(x, ps, st))), st


class GCNConv(AbstractGNNLayer):
    """x' = act(W (D^-1/2 Â D^-1/2 x) + b)  (layers.jl:147-239), CPU sparse-matmul semantics of GNN.jl."""

    def __init__(self, in_chs, out_chs=None, activation="identity", *, initialgraph=initialgraph, init_weight=None,
                 init_bias=zeros32, bias: bool = True, add_self_loops: bool = True, use_edge_weight: bool = False):
        if isinstance(in_chs, tuple):  # GCNConv(in => out, act): defaults to glorot_uniform (layers.jl:192-194)
            if out_chs is not None:
                activation = out_chs
            in_chs, out_chs = in_chs
            default_init = glorot_uniform
        else:  # GCNConv(in, out, act): defaults to glorot_normal (layers.jl:177-179)
            default_init = glorot_normal
        self.in_chs, self.out_chs = int(in_chs), int(out_chs)
        self.activation = _act_name(activation)  # NNlib.fast_act (layers.jl:180)
        self.initialgraph = wrapgraph(initialgraph)
        self.init_weight = init_weight or default_init
        self.init_bias = init_bias
        self.bias = bool(bias)
        self.add_self_loops = bool(add_self_loops)
        self.use_edge_weight = bool(use_edge_weight)

    def initialparameters(self, rng, device="cpu") -> NT:
        p = NT(weight=julia_array(self.init_weight(rng, self.out_chs, self.in_chs), device))
        if self.bias:
            dict.__setitem__(p, "bias", julia_array(self.init_bias(rng, self.out_chs, 1), device))
        return p

    def parameterlength(self) -> int:  # layers.jl:173-175
        return self.out_chs * (self.in_chs + 1) if self.bias else self.out_chs * self.in_chs

    def __call__(self, x: Tensor, ps, st: NT, edge_weight: Optional[Tensor] = None):
        g: GNNGraph = st.graph
        dev = x.device
        if edge_weight is not None and edge_weight.numel() != g.num_edges:  # layers.jl:207
            raise AssertionError(f"Wrong number of edge weights (expected {g.num_edges} but given {edge_weight.numel()})")
        if not self.bias:
            # the reference reads ps.bias unconditionally (layers.jl:238), so bias=false fails there with a field error
            raise AttributeError("GCNConv{bias=false}: the reference's call method reads ps.bias (layers.jl:238)")
        x_rm = rowmajor(x)
        _check_nodes(x_rm, g)
        if x_rm.shape[1] != self.in_chs:
            raise ValueError(f"DimensionMismatch: GCNConv({self.in_chs} => {self.out_chs}) got {x_rm.shape[1]} features")
        desc = _lib.GcnDesc(self.in_chs, self.out_chs, _lib.ACT[self.activation], 1, int(self.add_self_loops),
                            int(self.use_edge_weight))
        params = flat_params(ps, self.parameterlength())
        # the graph's own stored weights always go down: `degree(g, T; dir=:in, edge_weight)` with edge_weight === nothing
        # resolves to them (GNN.jl `_get_edge_weight`), whether or not use_edge_weight lets them scale the messages
        gw = g.w.to(dev) if g.w is not None else None
        ew = None if edge_weight is None else edge_weight.to(dev)
        y = ops.GcnFunction.apply(x_rm, params, g.handle(dev), desc, ew, gw)
        return from_rowmajor(y), st

    def __repr__(self):
        a = "" if self.activation == "identity" else f", {self.activation}"
        return f"GCNConv({self.in_chs} => {self.out_chs}{a})"


class SpectralConv(AbstractGNNLayer):
    """Fourier differentiation of a periodic function sampled on 2 pi j / n, j = 1..n (layers.jl:549-662):
    u'_i = 1/2 sum_j cos((x_i - x_j) n / 2) cot((x_i - x_j) / 2) u_j as `propagate(message, g, +)` over the complete digraph
    whose `edata.e` is x[t] - x[s].  No parameters.  The per-edge coefficient `cos(e n / 2) cot(e / 2) / 2` is static: it is
    evaluated once per graph in the precision `e` is stored in (Float64, as the reference builds it, layers.jl:642-646),
    rounded to float32 and cached; the call is one pass of the ordered aggregate kernel (and, in the backward, one pass over
    the transposed graph).  The reference would promote a Float32 input to Float64 through the Float64 `e`; this path
    computes in float32."""

    def __init__(self, n: int):
        self.n = int(n)
        self.initialgraph = None

    def initialstates(self, rng) -> NT:
        n = self.n
        ii = torch.arange(n).repeat_interleave(n)
        jj = torch.arange(n).repeat(n)
        keep = ii != jj
        s, t = ii[keep], jj[keep]  # Graphs.edges(complete_digraph(n)): source-major, ascending target
        x = torch.linspace(0.0, 2.0 * math.pi, n + 1, dtype=torch.float64)[1:]
        return NT(graph=GNNGraph(s, t, num_nodes=n, edata=(x[t] - x[s]).reshape(1, -1)))

    def initialparameters(self, rng, device="cpu") -> NT:
        return NT()

    def parameterlength(self) -> int:
        return 0

    def _coef(self, g: GNNGraph, dev) -> Tensor:
        e = g.edata["e"]
        key = ("spectral_coef", self.n, e.data_ptr(), e._version, str(dev))
        hit = g._cache.get("spectral_coef")
        if hit is None or hit[0] != key or hit[1] is not e:
            ed = e.reshape(-1).to(torch.float64)
            c = torch.cos(ed * self.n / 2) * (1.0 / torch.tan(ed / 2)) / 2  # layers.jl:654
            hit = (key, e, c.to(torch.float32).to(dev).contiguous())
            g._cache["spectral_coef"] = hit
        return hit[2]

    def __call__(self, x: Tensor, ps, st):
        g: GNNGraph = st["graph"]
        vec = x.dim() == 1  # layers.jl:659-662
        x2 = x.reshape(1, -1) if vec else x
        x_rm = rowmajor(x2)
        _check_nodes(x_rm, g)
        y = from_rowmajor(ops.WeightedSumFunction.apply(x_rm, g, self._coef(g, x.device)))
        return (y.reshape(-1) if vec else y), st

    def __repr__(self):
        return f"SpectralConv({self.n})"


def propagate_copy_xj(g: GNNGraph, aggr, x: Tensor, e: Optional[Tensor] = None) -> Tensor:
    """propagate(copy_xj | e_mul_xj, g, aggr; xj = x [, e]) on the ordered, atomic-free scatter path (forward only).
    This is the primitive the reference's SpectralConv known-answer test exercises (layers.jl:656)."""
    out = ops.aggregate(g.handle(x.device), _aggr_name(aggr), rowmajor(x), None if e is None else e.reshape(-1))
    return from_rowmajor(out)
