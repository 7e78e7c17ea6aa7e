"""ngpde-b200: B200-native message-passing hot path of NeuralGraphPDE.jl behind the Lux layer API.

Exports mirror /root/reference/src/NeuralGraphPDE.jl:27-33 (layer types, `updategraph`) plus the pieces of
GraphNeuralNetworks / Lux the reference re-exports or its tests use (`GNNGraph`, `rand_graph`, `batch`, `Dense`,
`Chain`, `setup`).  Importing the package does not need a GPU; calling a layer does (there is no CPU fallback).
"""
from . import _lib
from ._lib import NgpdeError
from .graph import GNNGraph, add_self_loops, batch, copy, from_rowmajor, rand_graph, rowmajor
from .layers import (AbstractGNNContainerLayer, AbstractGNNLayer, ExplicitEdgeConv, GCNConv, GNOConv, MPPDEConv,
                     SpectralConv, VMHConv, initialgraph, propagate_copy_xj, wrapgraph)
from .lux import (NT, Chain, ComponentArray, Dense, flat_params, glorot_normal, glorot_uniform, julia_array, merge,
                  ones32, setup, zeros32)
from .utils import drop, updategraph
from . import distributed, losses, optim, partition

__all__ = [
    "AbstractGNNLayer", "AbstractGNNContainerLayer", "ExplicitEdgeConv", "GCNConv", "VMHConv", "MPPDEConv", "GNOConv",
    "SpectralConv", "updategraph", "GNNGraph", "rand_graph", "batch", "add_self_loops", "copy", "Dense", "Chain", "setup", "NT",
    "ComponentArray", "NgpdeError",
]
