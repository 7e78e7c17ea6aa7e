# Persistent fixed-step Runge-Kutta integrator for  du/dt = ExplicitEdgeConv(u, ps, st)[1]  on graphs small enough that one
# right-hand side is launch latency (the reference's tutorial loop `solve(prob, Tsit5(); adaptive = false, dt)`,
# docs/src/tutorials/graph_node.md:53-66): ONE kernel launch integrates all steps (`ngpde_edgeconv_ode_forward`), a second one
# is its discrete adjoint (`ngpde_edgeconv_ode_adjoint`).  UNTESTED (no Julia toolchain in the build image); the same
# symbols are exercised by neuralgraphpde.jl_b200/ode.py (PersistentRK) and tests/test_gpu_ode.py.

struct RkTableau
    n_stages::Int32
    a::NTuple{64, Float32}    # a[i][j], row-major 8 x 8, used for j < i
    b::NTuple{8, Float32}
end

function RkTableau(A::AbstractVector, b::AbstractVector)
    S = length(b)
    S <= 8 || throw(ArgumentError("at most 8 stages"))
    a = zeros(Float32, 64)
    for i in 1:S, j in 1:(i - 1)
        a[(i - 1) * 8 + j] = Float32(A[i][j])
    end
    bb = zeros(Float32, 8)
    bb[1:S] .= Float32.(b)
    return RkTableau(Int32(S), Tuple(a), Tuple(bb))
end

const RK4 = RkTableau([Float32[], [0.5f0], [0.0f0, 0.5f0], [0.0f0, 0.0f0, 1.0f0]], [1 / 6, 1 / 3, 1 / 3, 1 / 6])
# Tsit5 (Tsitouras 2011), the b-row of the 5th-order solution; the 7th (FSAL) stage is not needed at fixed step
const TSIT5 = RkTableau(
    [Float64[], [0.161], [-0.008480655492356989, 0.335480655492357],
     [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
     [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
     [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383]],
    [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774])

function edgeconv_desc(l::ExplicitEdgeConv, x::CuMatrix{Float32}, st::NamedTuple)
    s = st.graph.ndata
    hs_keys = filter(!=(:x), ksym(s))
    snode = packed(s, (hs_keys..., :x))
    dpos = size(s.x, 1)
    desc = ConvDesc(FAM_EDGECONV, aggr_code(l.aggr), size(x, 1), size(snode, 1) - dpos, dpos, 0, 0, 0, 0, Mlp(l.ϕ), Mlp())
    return desc, snode
end

"""
    solve_persistent(l::ExplicitEdgeConv, u0, ps, st; dt, nsteps, tableau = TSIT5) -> (uT, traj)

`uT (dx, N)` after `nsteps` steps of size `dt`; `traj` holds every stage input (the adjoint's checkpoint).
Limits (checked by the library): aggr `+` / `mean`, dx <= 4, ϕ input <= 16, ϕ layers <= 32 wide, <= 4 layers.
"""
solve_persistent(l::ExplicitEdgeConv, u0::CuMatrix{Float32}, ps, st::NamedTuple; dt::Real, nsteps::Integer, tableau::RkTableau = TSIT5) =
    _solve_persistent(l, u0, flat(ps), st, Float32(dt), Int(nsteps), tableau)   # Zygote differentiates `flat`; the rrule below sees the flat vector

function _solve_persistent(l::ExplicitEdgeConv, u0::CuMatrix{Float32}, phi::CuVector{Float32}, st::NamedTuple, dt::Float32, nsteps::Int, tableau::RkTableau)
    h = handle(st.graph)
    desc, snode = edgeconv_desc(l, u0, st)
    u = copy(u0)
    traj = CuArray{Float32}(undef, size(u0, 1), size(u0, 2), Int(tableau.n_stages), Int(nsteps))
    rd, rt = Ref(desc), Ref(tableau)
    nb = GC.@preserve rd rt ccall((:ngpde_edgeconv_ode_workspace_bytes, libngpde), Csize_t, (Ptr{Cvoid}, Ptr{ConvDesc}, Ptr{RkTableau}), h.ptr, rd, rt)
    nb == 0 && check(-1)
    ws = workspace(nb)
    GC.@preserve h rd rt phi snode u traj ws check(ccall((:ngpde_edgeconv_ode_forward, libngpde), Cint,
        (Ptr{Cvoid}, Ptr{ConvDesc}, Ptr{RkTableau}, Float32, Int32, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
         CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
        h.ptr, rd, rt, Float32(dt), Int32(nsteps), pointer(phi), ptr(snode), pointer(u), pointer(traj), pointer(ws), length(ws), cuda_stream()))
    return u, traj
end

function ChainRulesCore.rrule(::typeof(_solve_persistent), l::ExplicitEdgeConv, u0::CuMatrix{Float32}, phi::CuVector{Float32}, st::NamedTuple,
                              dt::Float32, nsteps::Int, tableau::RkTableau)
    uT, traj = _solve_persistent(l, u0, phi, st, dt, nsteps, tableau)
    function solve_persistent_pullback(Δ)
        duT = unthunk(Δ[1])
        duT isa AbstractZero && return ntuple(_ -> NoTangent(), 8)
        h = handle(st.graph)
        desc, snode = edgeconv_desc(l, u0, st)
        lam = CuMatrix{Float32}(duT)           # in: dL/du(T); out: dL/du(0)
        dphi = similar(phi)
        rd, rt = Ref(desc), Ref(tableau)
        nb = GC.@preserve rd rt ccall((:ngpde_edgeconv_ode_workspace_bytes, libngpde), Csize_t, (Ptr{Cvoid}, Ptr{ConvDesc}, Ptr{RkTableau}), h.ptr, rd, rt)
        ws = workspace(nb)
        GC.@preserve h rd rt phi snode traj lam dphi ws check(ccall((:ngpde_edgeconv_ode_adjoint, libngpde), Cint,
            (Ptr{Cvoid}, Ptr{ConvDesc}, Ptr{RkTableau}, Float32, Int32, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32}, CuPtr{Float32},
             CuPtr{Float32}, CuPtr{Cvoid}, Csize_t, Ptr{Cvoid}),
            h.ptr, rd, rt, Float32(dt), Int32(nsteps), pointer(phi), ptr(snode), pointer(traj), pointer(lam), pointer(dphi), pointer(ws),
            length(ws), cuda_stream()))
        return NoTangent(), NoTangent(), lam, dphi, NoTangent(), NoTangent(), NoTangent(), NoTangent()
    end
    return (uT, traj), solve_persistent_pullback
end
