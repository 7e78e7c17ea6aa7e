"""Host-side logic of the drop-in mirror (no GPU): parameter/state trees, updategraph semantics
(/root/reference/test/runtests.jl:16-54, 166-206), ComponentArray ordering, and the C-ABI export list."""
import os
import re

import numpy as np
import pytest
import torch

import ngpde
from ngpde import (NT, Chain, ComponentArray, Dense, ExplicitEdgeConv, GCNConv, GNNGraph, GNOConv, MPPDEConv, VMHConv,
                   rand_graph, setup, updategraph)

from ngpde.lux import nfkc


def names(*ks):
    return [nfkc(k) for k in ks]


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def toy():
    return GNNGraph([0, 0, 1, 2], [1, 2, 0, 0])


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ngpde.h")).read()
    declared = set(re.findall(r"\b(ngpde_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = ngpde._lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libngpde.so does not export {name}"
    assert set(ngpde._lib.EXPORTS) == declared
    assert lib.ngpde_version() == 200


def test_enum_values_agree_between_header_python_and_julia_bindings():
    """The aggregation / option codes are part of the C ABI: include/ngpde.h is the source, the ctypes mirror (_lib.py) and the
    Julia binding (julia/NeuralGraphPDEB200/src/capi.jl) must carry the same numbers."""
    header = open(os.path.join(ROOT, "include", "ngpde.h")).read()
    vals = {k: int(v) for k, v in re.findall(r"\b(NGPDE_(?:AGGR|OPT)_[A-Z_]+)\s*=\s*(\d+)", header)}
    assert vals["NGPDE_AGGR_SUM"] == 0 and vals["NGPDE_AGGR_PROD"] == 4
    L = ngpde._lib
    assert L.AGGR == {"+": vals["NGPDE_AGGR_SUM"], "sum": vals["NGPDE_AGGR_SUM"], "mean": vals["NGPDE_AGGR_MEAN"],
                      "max": vals["NGPDE_AGGR_MAX"], "min": vals["NGPDE_AGGR_MIN"], "*": vals["NGPDE_AGGR_PROD"],
                      "prod": vals["NGPDE_AGGR_PROD"]}
    assert (L.OPT_TENSOR_CORES, L.OPT_GNO_FACTORED, L.OPT_HOIST, L.OPT_LAYERED, L.OPT_GNO_LAYERED) == tuple(
        vals[k] for k in ("NGPDE_OPT_TENSOR_CORES", "NGPDE_OPT_GNO_FACTORED", "NGPDE_OPT_HOIST", "NGPDE_OPT_LAYERED",
                          "NGPDE_OPT_GNO_LAYERED"))
    jl = open(os.path.join(ROOT, "julia", "NeuralGraphPDEB200", "src", "capi.jl")).read()
    assert "AGGR_SUM, AGGR_MEAN, AGGR_MAX, AGGR_MIN, AGGR_PROD = Int32.(0:4)" in jl
    m = re.search(r"const OPT_TENSOR_CORES, OPT_GNO_FACTORED, OPT_DEBUG_SKIP, OPT_HOIST, OPT_LAYERED, OPT_GNO_LAYERED = Int32\.\((\d):(\d)\)", jl)
    assert m and (int(m.group(1)), int(m.group(2))) == (0, 5)


def test_gcn_state_and_params():  # runtests.jl:16-25
    g = toy()
    l = GCNConv((3, 5), initialgraph=g)
    ps, st = setup(0, l)
    assert st == NT(graph=g)
    assert ps.weight.shape == (5, 3) and ps.bias.shape == (5, 1)
    assert l.parameterlength() == 5 * 4 and l.statelength() == 1
    assert ps.weight.T.is_contiguous()  # Julia column-major


def test_container_layers_state_trees():  # runtests.jl:27-54
    g = toy()
    gh = GNNGraph(g, ndata={"x": torch.rand(3, 3)})
    l = ExplicitEdgeConv(Dense(11, 5), initialgraph=gh)
    ps, st = setup(0, l)
    assert st == NT(ϕ=NT(), graph=gh)
    assert list(ps.keys()) == ["weight", "bias"]  # single sub-layer: parameters un-nested (devdoc.md:74-88)
    l = VMHConv(Dense(11, 5), Dense(9, 7), initialgraph=gh)
    ps, st = setup(0, l)
    assert st == NT(ϕ=NT(), γ=NT(), graph=gh)
    assert list(ps.keys()) == names("ϕ", "γ") and ps.ϕ.weight.shape == (5, 11) and ps.γ.weight.shape == (7, 9)
    l = MPPDEConv(Dense(19, 5), Dense(14, 7), initialgraph=gh)
    ps, st = setup(0, l)
    assert list(st.keys()) == names("ϕ", "ψ", "graph") and st.graph == gh
    l = GNOConv((5, 7), Dense(10, 35), initialgraph=gh)
    ps, st = setup(0, l)
    assert list(ps.keys()) == names("linear", "ϕ") and ps.linear.weight.shape == (7, 5)
    assert list(st.keys()) == names("linear", "ϕ", "graph")
    assert l.statelength() == 1 and l.parameterlength() == 7 * 6 + 35 * 11


def test_default_initialgraph_is_empty():
    ps, st = setup(0, GCNConv((3, 5)))
    assert st.graph.num_nodes == 0 and st.graph.num_edges == 0


def test_updategraph_identity_semantics():  # runtests.jl:166-185
    g = rand_graph(5, 4, bidirected=False, seed=0)
    l = GCNConv((3, 5), initialgraph=g)
    ps, st = setup(0, l)
    new_g = rand_graph(5, 7, bidirected=False, seed=1)
    new_st = updategraph(st, new_g)
    assert new_st.graph is new_g
    model = Chain(GCNConv((3, 5), initialgraph=g), GCNConv((5, 5), initialgraph=g))
    ps, st = setup(0, model)
    new_st = updategraph(st, new_g)
    assert new_st.layer_1.graph is new_st.layer_2.graph is new_g


def test_updategraph_data_only():  # runtests.jl:187-205
    g = rand_graph(5, 4, bidirected=False, seed=0)
    ps, st = setup(0, GCNConv((3, 5), initialgraph=g))
    ndata = torch.rand(3, g.num_nodes)
    new_st = updategraph(st, ndata=ndata)
    assert new_st.graph.ndata["x"] is ndata
    assert new_st.graph._topo is st.graph._topo  # topology (and its cached CSR handle) is shared
    model = Chain(GCNConv((3, 5), initialgraph=g), GCNConv((5, 5), initialgraph=g))
    ps, st = setup(0, model)
    new_st = updategraph(st, ndata=ndata)
    assert new_st.layer_1.graph.ndata["x"] is new_st.layer_2.graph.ndata["x"] is ndata
    assert updategraph(NT(), g) == NT()


def test_graph_equality_and_copy():
    g = toy()
    assert ngpde.copy(g) == g and ngpde.copy(g) is not g
    assert GNNGraph([0, 0, 1, 2], [1, 2, 0, 0]) == g
    assert GNNGraph([0, 0, 1, 2], [1, 2, 0, 1]) != g
    gh = GNNGraph(g, ndata=torch.rand(3, 3))
    assert gh != g and list(gh.ndata) == ["x"]
    ge = GNNGraph(g, edata=torch.rand(2, 4))
    assert list(ge.edata) == ["e"]
    gb = ngpde.batch([GNNGraph(g, gdata={"θ": torch.rand(4)}), GNNGraph(g, gdata={"θ": torch.rand(4)})])
    assert gb.num_nodes == 6 and gb.num_edges == 8 and gb.num_graphs == 2 and gb.gdata["θ"].shape == (4, 2)
    assert torch.equal(gb.s, torch.tensor([0, 0, 1, 2, 3, 3, 4, 5]))
    gl = ngpde.add_self_loops(g)
    assert torch.equal(gl.s[-3:], torch.arange(3)) and torch.equal(gl.t[-3:], torch.arange(3))
    g1 = GNNGraph([1, 1, 2, 3], [2, 3, 1, 1], index_base=1)
    assert g1 == g


def test_component_array_order_and_views():
    l = VMHConv(Chain(Dense(6, 4, "tanh"), Dense(4, 3)), Dense(5, 2), initialgraph=toy())
    ps, _ = setup(0, l)
    ca = ComponentArray(ps)
    assert len(ca) == l.parameterlength()
    # field order, each array column-major: ϕ.layer_1.weight, ϕ.layer_1.bias, ϕ.layer_2.weight, ...
    w1 = ps.ϕ.layer_1.weight
    assert torch.equal(ca.data[:24], w1.T.reshape(-1))
    assert torch.equal(ca.data[24:28], ps.ϕ.layer_1.bias.reshape(-1))
    assert torch.equal(ca.ϕ.layer_2.weight, ps.ϕ.layer_2.weight)
    # sub-trees are zero-copy views of the flat vector
    assert ca.ϕ.data.data_ptr() == ca.data.data_ptr()
    assert ca.γ.data.data_ptr() == ca.data.data_ptr() + 4 * (24 + 4 + 12 + 3)
    assert ngpde.flat_params(ca.γ, 12).data_ptr() == ca.γ.data.data_ptr()
    flat = ngpde.flat_params(ps.ϕ, 43)
    assert torch.equal(flat, ca.ϕ.data)
    with pytest.raises(ValueError):
        ngpde.flat_params(ps.ϕ, 44)


def test_no_cpu_fallback():
    g = GNNGraph(toy(), ndata={"x": torch.rand(3, 3)})
    l = VMHConv(Dense(11, 5), Dense(9, 7), initialgraph=g)
    ps, st = setup(0, l)
    with pytest.raises(ngpde.NgpdeError):
        l(torch.randn(4, 3), ps, st)


def test_gcn_argument_checks():
    g = toy()
    l = GCNConv((3, 5), initialgraph=g)
    ps, st = setup(0, l)
    with pytest.raises(AssertionError, match="Wrong number of edge weights"):  # layers.jl:207
        l(torch.randn(3, 3), ps, st, torch.ones(3))


def test_unsupported_pieces_fail_loudly():
    with pytest.raises(ValueError):
        Dense(3, 4, "mish")
    with pytest.raises(ValueError):
        VMHConv(Dense(3, 4), Dense(3, 4), aggr="median")
