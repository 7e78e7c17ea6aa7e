# The training step either side of the adjoint, on the device (SURVEY.md section 8f-4): the optimiser updates and losses
# the reference's tutorials use (docs/src/tutorials/graph_node.md:100-129, VMH.md:97-109,140-143) as fused kernels over
# the flat parameter vector.  State layout follows Optimisers.jl (`(mt, vt, βt)` for Adam, `(g, η)` for Rprop), so these
# can sit behind `Optimisers.apply!` methods specialised on CuVector{Float32} if wanted.

"""
    adam_step!(x, g, m, v, βt; eta = 0.001f0, beta = (0.9f0, 0.999f0), epsilon = 1f-8) -> βt .* beta

In place `x .-= m̂ / (sqrt(v̂) + ϵ) * η` with the Optimisers.jl operation order (bit-exact against its Float32 CPU path).
"""
function adam_step!(x::CuVector{Float32}, g::CuVector{Float32}, m::CuVector{Float32}, v::CuVector{Float32}, βt::NTuple{2, Float32};
                    eta = 0.001f0, beta = (0.9f0, 0.999f0), epsilon = 1.0f-8)
    GC.@preserve x g m v check(ccall((:ngpde_adam_step, libngpde), Cint,
        (CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Float32, Float32, Float32, Float32, Float32, Float32, Ptr{Cvoid}),
        pointer(x), pointer(g), pointer(m), pointer(v), length(x), eta, beta[1], beta[2], epsilon, βt[1], βt[2], cuda_stream()))
    return βt .* Float32.(beta)
end

"""
    rprop_step!(x, g, gprev, η; ell = (0.5f0, 1.2f0), gamma = (1f-6, 50f0))
"""
function rprop_step!(x::CuVector{Float32}, g::CuVector{Float32}, gprev::CuVector{Float32}, η::CuVector{Float32};
                     ell = (0.5f0, 1.2f0), gamma = (1.0f-6, 50.0f0))
    GC.@preserve x g gprev η check(ccall((:ngpde_rprop_step, libngpde), Cint,
        (CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, Int64, Float32, Float32, Float32, Float32, Ptr{Cvoid}),
        pointer(x), pointer(g), pointer(gprev), pointer(η), length(x), ell[1], ell[2], gamma[1], gamma[2], cuda_stream()))
    return nothing
end

loss_workspace() = CuArray{UInt8}(undef, Int(ccall((:ngpde_loss_workspace_bytes, libngpde), Csize_t, ())))

"""
    mse_loss(ŷ, y) -> (loss::Float32, dŷ)       loss = mean(abs2, ŷ .- y), dŷ = 2 (ŷ - y) / length(y)
"""
function mse_loss(ŷ::CuArray{Float32}, y::CuArray{Float32})
    size(ŷ) == size(y) || throw(DimensionMismatch("mse: $(size(ŷ)) vs $(size(y))"))
    loss, dŷ, ws = CUDA.zeros(Float32, 1), similar(ŷ), loss_workspace()
    GC.@preserve ŷ y loss dŷ ws check(ccall((:ngpde_mse_loss, libngpde), Cint,
        (CuPtr{Float32}, CuPtr{Float32}, Int64, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
        pointer(ŷ), pointer(y), length(y), pointer(loss), pointer(dŷ), pointer(ws), length(ws), cuda_stream()))
    return loss, dŷ
end
mse(ŷ::CuArray{Float32}, y::CuArray{Float32}) = CUDA.@allowscalar mse_loss(ŷ, y)[1][1]
function ChainRulesCore.rrule(::typeof(mse), ŷ::CuArray{Float32}, y::CuArray{Float32})
    loss, dŷ = mse_loss(ŷ, y)
    mse_pullback(Δ) = (NoTangent(), dŷ .* Float32(unthunk(Δ)), NoTangent())
    return CUDA.@allowscalar(loss[1]), mse_pullback
end

"""
    logitcrossentropy_loss(ŷ, y, mask = nothing) -> (loss, dŷ)

`mean(-sum(y .* logsoftmax(ŷ[:, mask]); dims = 1))` (graph_node.md:100-106); `mask` is a vector of 1-based column indices.
"""
function logitcrossentropy_loss(ŷ::CuMatrix{Float32}, y::CuMatrix{Float32}, mask::Union{Nothing, AbstractVector{<:Integer}} = nothing)
    C, N = size(ŷ)
    idx = mask === nothing ? nothing : CuVector{Int32}(Int32.(mask) .- Int32(1))
    nm = mask === nothing ? N : length(mask)
    size(y) == (C, nm) || throw(DimensionMismatch("logitcrossentropy: y is $(size(y)), expected ($C, $nm)"))
    loss, dŷ, ws = CUDA.zeros(Float32, 1), similar(ŷ), loss_workspace()
    GC.@preserve ŷ y idx loss dŷ ws check(ccall((:ngpde_logit_cross_entropy, libngpde), Cint,
        (CuPtr{Float32}, Int64, Int32, CuPtr{Float32}, CuPtr{Int32}, Int64, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
        pointer(ŷ), N, C, pointer(y), idx === nothing ? CU_NULL : pointer(idx), nm, pointer(loss), pointer(dŷ), pointer(ws), length(ws),
        cuda_stream()))
    return loss, dŷ
end
