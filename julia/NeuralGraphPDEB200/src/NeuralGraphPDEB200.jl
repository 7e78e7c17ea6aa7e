"""
    NeuralGraphPDEB200

Drop-in B200 back end for the message-passing hot path of NeuralGraphPDE.jl.  Loading this package next to
`NeuralGraphPDE` adds methods, specialised on `CuArray{Float32}` inputs, for the five layer calls

    (l::ExplicitEdgeConv)(x, ps, st)   src/layers.jl:94-112 of the reference
    (l::GCNConv)(x, ps, st[, w])       :200-239
    (l::VMHConv)(x, ps, st)            :308-332
    (l::MPPDEConv)(x, ps, st)          :390-422
    (l::GNOConv)(x, ps, st)            :509-547

that route to the fused sm_100a kernels of `libngpde.so` (C ABI: include/ngpde.h) through `ccall`, and ChainRules
`rrule`s that route Zygote's pullback to the hand-written backward kernels.  Layer types, constructors, parameter and
state NamedTuples, `updategraph` and the flat `ComponentArray(ps)` ordering are the reference's own and are untouched, so
the exported names are simply re-exported (`src/NeuralGraphPDE.jl:27-33`).

STATUS: UNTESTED.  The image this was written in has no Julia toolchain; the same exported symbols are exercised by the
Python/ctypes mirror (`neuralgraphpde.jl_b200/`), which is what `tests/` run.  `test/runtests.jl` is the acceptance test
to run on a machine that has Julia, the reference package and a B200.
"""
module NeuralGraphPDEB200

using CUDA
using ChainRulesCore
using ComponentArrays
using GraphNeuralNetworks
using GraphNeuralNetworks.GNNGraphs: GNNGraph, edge_index, get_edge_weight
using Lux
using NNlib
using Statistics: mean
using NeuralGraphPDE
import NeuralGraphPDE: AbstractGNNLayer, AbstractGNNContainerLayer, ExplicitEdgeConv, GCNConv, VMHConv, MPPDEConv,
                       GNOConv, SpectralConv, updategraph

include("capi.jl")
include("graphs.jl")
include("conv.jl")
include("layers.jl")
include("train.jl")
include("ode.jl")
include("dist.jl")

# the reference's export list (src/NeuralGraphPDE.jl:27-33), unchanged
export AbstractGNNLayer, AbstractGNNContainerLayer
export ExplicitEdgeConv, GCNConv, VMHConv, MPPDEConv, GNOConv, SpectralConv
export updategraph
# what this back end adds
export adam_step!, rprop_step!, mse_loss, logitcrossentropy_loss
export solve_persistent, RkTableau, RK4, TSIT5
export NodePartition, HaloExchange, partition_nodes, morton_order, halo_forward, halo_backward, allreduce_sum!

end # module
