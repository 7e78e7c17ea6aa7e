# Multi-GPU behind the same C ABI (SURVEY.md section 8e): the node partitioner, an NCCL communicator created from a unique id
# (one Julia process per GPU; ship the id over MPI.jl / Distributed / a file), and the per-RHS halo exchange with its rrule.

struct NodePartition
    ptr::Ptr{Cvoid}
    world::Int
    rank::Int
end

const PA = (bounds = 0, halo_global = 1, recv_counts = 2, send_counts = 3, send_local = 4, s_local = 5, t_local = 6,
            edge_ids = 7, seg_rows = 8, seg_ptr = 9, seg_pos = 10, peer_recv_offset = 11)

"""
    partition_nodes(g_or_(s, t, N), world, rank; by = :edges, bounds = nothing) -> NodePartition

`s`, `t`: HOST Int64 vectors of the full COO lists, 1-based as Julia stores them; `rank` is 0-based.
"""
function partition_nodes(s::Vector{Int64}, t::Vector{Int64}, num_nodes::Integer, world::Integer, rank::Integer; by = :edges, bounds = nothing)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    b = bounds === nothing ? C_NULL : pointer(bounds)
    GC.@preserve s t bounds check(ccall((:ngpde_partition_create, libngpde), Cint,
        (Ref{Ptr{Cvoid}}, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Int32, Int32, Int32, Int32, Int32, Ptr{Int64}),
        h, num_nodes, length(s), pointer(s), pointer(t), IDX_I64, 1, world, rank, by === :edges ? 1 : 0, b))
    return NodePartition(h[], world, rank)
end
destroy(p::NodePartition) = ccall((:ngpde_partition_destroy, libngpde), Cint, (Ptr{Cvoid},), p.ptr)

"0-based Int64 copy of one of the plan's arrays (`PA` names)."
function plan_array(p::NodePartition, which::Symbol)
    ptr_, n = Ref{Ptr{Int64}}(C_NULL), Ref{Int64}(0)
    check(ccall((:ngpde_partition_array, libngpde), Cint, (Ptr{Cvoid}, Int32, Ref{Ptr{Int64}}, Ref{Int64}), p.ptr, getfield(PA, which), ptr_, n))
    return n[] == 0 ? Int64[] : copy(unsafe_wrap(Array, ptr_[], n[]))
end

"Z-order permutation of the nodes from `pos (dim, N)` (host Float32): `order[k]` = 1-based id of the k-th node on the curve."
function morton_order(pos::Matrix{Float32})
    dim, n = size(pos)
    order = Vector{Int64}(undef, n)
    GC.@preserve pos order check(ccall((:ngpde_morton_order, libngpde), Cint, (Ptr{Float32}, Int64, Int32, Ptr{Int64}), pointer(pos), n, dim, pointer(order)))
    return order .+ 1
end

mutable struct Comm
    ptr::Ptr{Cvoid}
end
unique_id() = (id = zeros(UInt8, 128); check(ccall((:ngpde_comm_unique_id, libngpde), Cint, (Ptr{UInt8}, Cstring), id, C_NULL)); id)
function Comm(id::Vector{UInt8}, world::Integer, rank::Integer)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ngpde_comm_init, libngpde), Cint, (Ref{Ptr{Cvoid}}, Ptr{UInt8}, Int32, Int32, Cstring), h, id, world, rank, C_NULL))
    c = Comm(h[])
    finalizer(x -> ccall((:ngpde_comm_destroy, libngpde), Cint, (Ptr{Cvoid},), x.ptr), c)
    return c
end
allreduce_sum!(c::Comm, buf::CuVector{Float32}) =
    (GC.@preserve buf check(ccall((:ngpde_allreduce_sum, libngpde), Cint, (Ptr{Cvoid}, CuPtr{Float32}, Int64, Ptr{Cvoid}), c.ptr, pointer(buf), length(buf), cuda_stream())); buf)

mutable struct HaloExchange
    ptr::Ptr{Cvoid}
    n_owned::Int
    n_halo::Int
    comm::Union{Nothing, Comm}
end
function HaloExchange(p::NodePartition, comm::Union{Nothing, Comm})
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ngpde_halo_create, libngpde), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), h, p.ptr,
                comm === nothing ? C_NULL : comm.ptr, cuda_stream()))
    b = plan_array(p, :bounds)
    ex = HaloExchange(h[], Int(b[p.rank + 2] - b[p.rank + 1]), length(plan_array(p, :halo_global)), comm)
    finalizer(x -> ccall((:ngpde_halo_destroy, libngpde), Cint, (Ptr{Cvoid},), x.ptr), ex)
    return ex
end

"`x_owned (d, n_owned)` -> `x_local (d, n_owned + n_halo)`: boundary rows exchanged over NCCL inside the library."
function halo_forward(ex::HaloExchange, x_owned::CuMatrix{Float32})
    d = size(x_owned, 1)
    x_local = CuMatrix{Float32}(undef, d, ex.n_owned + ex.n_halo)
    GC.@preserve x_owned x_local check(ccall((:ngpde_halo_forward, libngpde), Cint, (Ptr{Cvoid}, CuPtr{Float32}, Int32, CuPtr{Float32}, Ptr{Cvoid}),
        ex.ptr, pointer(x_owned), d, pointer(x_local), cuda_stream()))
    return x_local
end
"The transpose: halo cotangents go home and are added per owned row in fixed peer order (deterministic)."
function halo_backward(ex::HaloExchange, dx_local::CuMatrix{Float32})
    d = size(dx_local, 1)
    dx_owned = CuMatrix{Float32}(undef, d, ex.n_owned)
    GC.@preserve dx_local dx_owned check(ccall((:ngpde_halo_backward, libngpde), Cint, (Ptr{Cvoid}, CuPtr{Float32}, Int32, CuPtr{Float32}, Ptr{Cvoid}),
        ex.ptr, pointer(dx_local), d, pointer(dx_owned), cuda_stream()))
    return dx_owned
end
function ChainRulesCore.rrule(::typeof(halo_forward), ex::HaloExchange, x_owned::CuMatrix{Float32})
    halo_pullback(Δ) = (NoTangent(), NoTangent(), halo_backward(ex, CuMatrix{Float32}(unthunk(Δ))))
    return halo_forward(ex, x_owned), halo_pullback
end
