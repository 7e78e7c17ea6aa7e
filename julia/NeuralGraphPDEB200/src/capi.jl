# Raw bindings of include/ngpde.h.  One Julia function per exported symbol, same argument order.  Every array argument is
# caller-owned device memory (`CuPtr`); callers wrap the ccall in `GC.@preserve`.

const libngpde = get(ENV, "NGPDE_LIB", joinpath(@__DIR__, "..", "..", "..", "neuralgraphpde.jl_b200", "libngpde.so"))

struct NgpdeError <: Exception
    code::Int
    msg::String
end
Base.showerror(io::IO, e::NgpdeError) = print(io, "libngpde error ", e.code, ": ", e.msg)

last_error() = unsafe_string(ccall((:ngpde_last_error, libngpde), Cstring, ()))
@inline check(rc::Integer) = rc == 0 ? nothing : throw(NgpdeError(rc, last_error()))

cuda_stream() = Base.unsafe_convert(Ptr{Cvoid}, CUDA.stream().handle)   # calls are asynchronous on the task's stream

# ---- enums (include/ngpde.h) ----
const ACT_IDENTITY, ACT_RELU, ACT_TANH, ACT_SIGMOID, ACT_SWISH, ACT_GELU, ACT_SOFTPLUS, ACT_ELU, ACT_LEAKYRELU = Int32.(0:8)
const AGGR_SUM, AGGR_MEAN, AGGR_MAX, AGGR_MIN, AGGR_PROD = Int32.(0:4)
const FAM_EDGECONV, FAM_VMH, FAM_MPPDE, FAM_GNO = Int32.(0:3)
const IDX_I32, IDX_I64 = Int32(0), Int32(1)
const MAX_LAYERS = 8

# Lux resolves activations through NNlib.fast_act (tanh -> tanh_fast, ...); both spellings map to the same kernel branch
const ACT = IdDict{Any, Int32}(identity => ACT_IDENTITY, NNlib.relu => ACT_RELU, tanh => ACT_TANH,
                               NNlib.tanh_fast => ACT_TANH, NNlib.sigmoid => ACT_SIGMOID, NNlib.sigmoid_fast => ACT_SIGMOID,
                               NNlib.swish => ACT_SWISH, NNlib.gelu => ACT_GELU, NNlib.softplus => ACT_SOFTPLUS,
                               NNlib.elu => ACT_ELU, NNlib.leakyrelu => ACT_LEAKYRELU)
const AGGR = IdDict{Any, Int32}((+) => AGGR_SUM, mean => AGGR_MEAN, max => AGGR_MAX, min => AGGR_MIN, (*) => AGGR_PROD)

act_code(f) = get(ACT, f) do
    throw(ArgumentError("activation $f has no fused kernel branch; supported: $(collect(keys(ACT)))"))
end
aggr_code(f) = get(AGGR, f) do
    throw(ArgumentError("aggregation $f is not supported by the fused kernels (+, *, mean, max, min)"))
end

# ---- struct mirrors ----
struct Mlp
    n_layers::Int32
    dims::NTuple{MAX_LAYERS + 1, Int32}
    act::NTuple{MAX_LAYERS, Int32}
    has_bias::NTuple{MAX_LAYERS, Int32}
end
Mlp() = Mlp(0, ntuple(_ -> Int32(0), MAX_LAYERS + 1), ntuple(_ -> Int32(0), MAX_LAYERS), ntuple(_ -> Int32(0), MAX_LAYERS))

struct ConvDesc
    family::Int32
    aggr::Int32
    dx::Int32
    dhs::Int32
    dpos::Int32
    de::Int32
    dtheta::Int32
    gno_in::Int32
    gno_out::Int32
    phi::Mlp
    node::Mlp
end

struct ConvIO
    x::CuPtr{Float32}
    snode::CuPtr{Float32}
    edata::CuPtr{Float32}
    theta::CuPtr{Float32}
    phi_params::CuPtr{Float32}
    node_params::CuPtr{Float32}
    mbar::CuPtr{Float32}
    y::CuPtr{Float32}
    dy::CuPtr{Float32}
    dx::CuPtr{Float32}
    dphi_params::CuPtr{Float32}
    dnode_params::CuPtr{Float32}
    state::CuPtr{Cvoid}       # optional forward -> backward state (ngpde_conv_state_bytes); CU_NULL: the backward recomputes
end

struct GcnDesc
    in_chs::Int32
    out_chs::Int32
    act::Int32
    has_bias::Int32
    add_self_loops::Int32
    use_edge_weight::Int32
end

ptr(::Nothing) = CU_NULL
ptr(a::CuArray) = pointer(a)

# `Mlp` of a Lux Dense / Chain of Dense (layer widths, activation codes, bias flags)
dense_layers(d::Lux.Dense) = (d,)
dense_layers(c::Lux.Chain) = Tuple(values(c.layers))
function Mlp(model)
    ls = dense_layers(model)
    all(l -> l isa Lux.Dense, ls) || throw(ArgumentError("the fused kernels take a Dense or a Chain of Dense layers"))
    n = length(ls)
    n <= MAX_LAYERS || throw(ArgumentError("at most $MAX_LAYERS Dense layers per MLP"))
    dims = ntuple(i -> i == 1 ? Int32(ls[1].in_dims) : (i <= n + 1 ? Int32(ls[i - 1].out_dims) : Int32(0)), MAX_LAYERS + 1)
    act = ntuple(i -> i <= n ? act_code(ls[i].activation) : Int32(0), MAX_LAYERS)
    hb = ntuple(i -> i <= n ? Int32(hasbias(ls[i])) : Int32(0), MAX_LAYERS)
    return Mlp(Int32(n), dims, act, hb)
end
hasbias(::Lux.Dense{use_bias}) where {use_bias} = use_bias
outdim(model) = Int(dense_layers(model)[end].out_dims)

# ---- library ----
ngpde_version() = ccall((:ngpde_version, libngpde), Cint, ())
# process-wide switches (include/ngpde.h): all default to 1 except OPT_DEBUG_SKIP; OPT_LAYERED / OPT_GNO_LAYERED choose the
# GEMM-per-Dense-layer evaluation of wide MLPs / of the factored GNOConv (more workspace, see INTEGRATION.md)
const OPT_TENSOR_CORES, OPT_GNO_FACTORED, OPT_DEBUG_SKIP, OPT_HOIST, OPT_LAYERED, OPT_GNO_LAYERED = Int32.(0:5)
set_option(option::Integer, value::Integer) = check(ccall((:ngpde_set_option, libngpde), Cint, (Int32, Int32), option, value))

# ---- graph handle ----
function graph_create(num_nodes, num_edges, s, t; index_dtype = IDX_I64, index_base = 1, on_device = true, num_graphs = 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    P = on_device ? CuPtr{Cvoid} : Ptr{Cvoid}
    GC.@preserve s t check(ccall((:ngpde_graph_create, libngpde), Cint,
        (Ref{Ptr{Cvoid}}, Int64, Int64, P, P, Int32, Int32, Int32, Int64, Ptr{Cvoid}),
        h, num_nodes, num_edges, pointer(s), pointer(t), index_dtype, index_base, on_device ? 1 : 0, num_graphs, cuda_stream()))
    return h[]
end
graph_destroy(h::Ptr{Cvoid}) = ccall((:ngpde_graph_destroy, libngpde), Cint, (Ptr{Cvoid},), h)
graph_num_nodes(h) = ccall((:ngpde_graph_num_nodes, libngpde), Int64, (Ptr{Cvoid},), h)
graph_num_edges(h) = ccall((:ngpde_graph_num_edges, libngpde), Int64, (Ptr{Cvoid},), h)
function graph_array(h, which::Integer; with_self_loops = false)
    p, n = Ref{CuPtr{Cvoid}}(CU_NULL), Ref{Int64}(0)
    check(ccall((:ngpde_graph_array, libngpde), Cint, (Ptr{Cvoid}, Int32, Int32, Ref{CuPtr{Cvoid}}, Ref{Int64}), h, which, with_self_loops, p, n))
    out = CUDA.zeros(Int32, n[])
    GC.@preserve out check(ccall((:ngpde_graph_array_copy, libngpde), Cint, (Ptr{Cvoid}, Int32, Int32, CuPtr{Int32}, Int64, Ptr{Cvoid}),
        h, which, with_self_loops, pointer(out), n[], cuda_stream()))
    return out
end

# ---- layer calls ----
conv_workspace_bytes(h, desc::ConvDesc, backward::Bool) =
    ccall((:ngpde_conv_workspace_bytes, libngpde), Csize_t, (Ptr{Cvoid}, Ref{ConvDesc}, Int32), h, desc, backward)
conv_forward!(h, desc::ConvDesc, io::ConvIO, ws::CuArray{UInt8}) =
    check(ccall((:ngpde_conv_forward, libngpde), Cint, (Ptr{Cvoid}, Ref{ConvDesc}, Ref{ConvIO}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
        h, desc, io, pointer(ws), length(ws), cuda_stream()))
conv_backward!(h, desc::ConvDesc, io::ConvIO, ws::CuArray{UInt8}) =
    check(ccall((:ngpde_conv_backward, libngpde), Cint, (Ptr{Cvoid}, Ref{ConvDesc}, Ref{ConvIO}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
        h, desc, io, pointer(ws), length(ws), cuda_stream()))
gcn_workspace_bytes(h, desc::GcnDesc, backward::Bool) =
    ccall((:ngpde_gcn_workspace_bytes, libngpde), Csize_t, (Ptr{Cvoid}, Ref{GcnDesc}, Int32), h, desc, backward)
gcn_forward!(h, desc::GcnDesc, x, w, b, ew, gw, y, ws) =
    check(ccall((:ngpde_gcn_conv_forward, libngpde), Cint,
        (Ptr{Cvoid}, Ref{GcnDesc}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
         CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}), h, desc, x, w, b, ew, gw, y, pointer(ws), length(ws), cuda_stream()))
gcn_backward!(h, desc::GcnDesc, x, w, b, ew, gw, y, dy, dx, dw, db, ws) =
    check(ccall((:ngpde_gcn_conv_backward, libngpde), Cint,
        (Ptr{Cvoid}, Ref{GcnDesc}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
         CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
        h, desc, x, w, b, ew, gw, y, dy, dx, dw, db, pointer(ws), length(ws), cuda_stream()))
aggregate!(h, aggr, x, d, w, out) =
    check(ccall((:ngpde_aggregate, libngpde), Cint, (Ptr{Cvoid}, Int32, CuPtr{Float32}, Int32, CuPtr{Float32}, CuPtr{Float32}, Ptr{Cvoid}),
        h, aggr, x, d, w, out, cuda_stream()))
