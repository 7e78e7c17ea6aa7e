# Device layout cache.  The CSR layout (libngpde graph handle) hangs off the IDENTITY of the graph's COO source vector, not
# off `st`: `st == (graph = g,)` keeps holding after a call (test/runtests.jl:21,24 of the reference), shallow copies made
# by `updategraph(st; ndata = ...)` / `copy(g; kwargs...)` (src/utils.jl:8,24-31) share the vectors and therefore the
# handle, and a new topology brings new vectors and therefore a new handle.

mutable struct GraphHandle
    ptr::Ptr{Cvoid}
    function GraphHandle(p::Ptr{Cvoid})
        h = new(p)
        finalizer(h) do x
            x.ptr == C_NULL || graph_destroy(x.ptr)
            x.ptr = C_NULL
        end
        return h
    end
end
Base.unsafe_convert(::Type{Ptr{Cvoid}}, h::GraphHandle) = h.ptr

const HANDLES = WeakKeyDict{Any, GraphHandle}()
const HANDLES_LOCK = ReentrantLock()

"""
    handle(g::GNNGraph) -> GraphHandle

Stable dst-sorted CSR, edge permutation, src-sorted transpose, work-unit lists and the merged GCN adjacency, built once per
topology on the device (`ngpde_graph_create`).  Julia's 1-based Int64 COO vectors are passed as they are (`index_base = 1`).
"""
function handle(g::GNNGraph)
    s, t = edge_index(g)
    s isa CuArray || throw(ArgumentError("the B200 path needs the graph on the GPU (`g |> gpu`); there is no CPU fallback here"))
    lock(HANDLES_LOCK) do
        get!(HANDLES, s) do
            GraphHandle(graph_create(g.num_nodes, g.num_edges, s, t; index_dtype = eltype(s) == Int32 ? IDX_I32 : IDX_I64,
                                     index_base = 1, on_device = true, num_graphs = g.num_graphs))
        end
    end
end

# `[items][sum D]` row-major == Julia `(sum D, items)` column-major: a vcat of the fields in key order, Float32
function packed(nt::NamedTuple, keys_)
    isempty(keys_) && return nothing
    parts = map(k -> f32matrix(getfield(nt, k)), keys_)
    return length(parts) == 1 ? parts[1] : vcat(parts...)
end
f32matrix(a::CuMatrix{Float32}) = a
f32matrix(a::CuVector{Float32}) = reshape(a, 1, :)
function f32matrix(a::CuArray)
    # Float64 side data promotes the whole call in the reference (test/runtests.jl:58-61); this path computes in Float32
    @warn "graph data of eltype $(eltype(a)) is converted to Float32 on the B200 path" maxlog = 1
    return f32matrix(Float32.(a))
end
ChainRulesCore.@non_differentiable handle(::Any)
ChainRulesCore.@non_differentiable packed(::Any, ::Any)
