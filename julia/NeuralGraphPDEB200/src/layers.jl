# The five layer calls on GPU arrays.  Each builds the descriptor of include/ngpde.h for its family, packs the static graph
# data in the reference's `vcat` row order, and calls the fused primitive.  `st` is returned unchanged, as in the reference
# (the sub-layer states of Dense / Chain are empty NamedTuples).

const CuF32Mat = CuMatrix{Float32}

named(x::NamedTuple) = x
named(x::AbstractArray) = (; preservedname = x)
drop_x(nt::NamedTuple) = Base.structdiff(nt, NamedTuple{(:x,)})
ksym(nt::NamedTuple) = collect(keys(nt))

function check_no_collision(x::NamedTuple, s::NamedTuple, who)
    both = intersect(keys(x), keys(s))
    isempty(both) || throw(ArgumentError("$who: field(s) $both appear both in the input and in st.graph.ndata; merge(x, ndata) " *
                                         "would replace the trainable input by static data (layers.jl:110,324) -- rename one"))
end

# ---- ExplicitEdgeConv: m = phi([h_i; h_j; pos_j - pos_i]), y = aggr(m)        reference src/layers.jl:94-112 ----
function (l::ExplicitEdgeConv)(x::CuF32Mat, ps, st::NamedTuple)
    y, st = l((; preservedname = x), ps, st)
    return y, st
end
function (l::ExplicitEdgeConv)(x::NamedTuple{K, <:Tuple{Vararg{CuF32Mat}}}, ps, st::NamedTuple) where {K}
    g = st.graph
    s = g.ndata
    haskey(s, :x) || throw(ArgumentError("ExplicitEdgeConv needs the coordinates in st.graph.ndata.x"))
    check_no_collision(x, s, "ExplicitEdgeConv")
    h = length(x) == 1 ? first(values(drop_x(x))) : vcat(values(drop_x(x))...)
    hs_keys = filter(!=(:x), ksym(s))
    snode = packed(s, (hs_keys..., :x))
    dpos = size(s.x, 1)
    desc = ConvDesc(FAM_EDGECONV, aggr_code(l.aggr), size(h, 1), size(snode, 1) - dpos, dpos, 0, 0, 0, 0, Mlp(l.ϕ), Mlp())
    y, _ = fused_conv(handle(g), desc, h, snode, nothing, nothing, flat(ps), nothing)
    return y, st
end

# ---- VMHConv: m = phi([h_i; h_j - h_i; pos_j - pos_i]); y = gamma([x; aggr(m)])   reference src/layers.jl:308-332 ----
function (l::VMHConv)(x::CuF32Mat, ps, st::NamedTuple)
    y, st = l((; preservedname = x), ps, st)
    return y, st
end
function (l::VMHConv)(x::NamedTuple{K, <:Tuple{Vararg{CuF32Mat}}}, ps, st::NamedTuple) where {K}
    g = st.graph
    s = g.ndata
    haskey(s, :x) || throw(ArgumentError("VMHConv needs the coordinates in st.graph.ndata.x"))
    check_no_collision(x, s, "VMHConv")
    h = length(x) == 1 ? first(values(x)) : vcat(values(x)...)
    hs_keys = filter(!=(:x), ksym(s))
    snode = packed(s, (hs_keys..., :x))
    dpos = size(s.x, 1)
    desc = ConvDesc(FAM_VMH, aggr_code(l.aggr), size(h, 1), size(snode, 1) - dpos, dpos, 0, 0, 0, 0, Mlp(l.ϕ), Mlp(l.γ))
    y, _ = fused_conv(handle(g), desc, h, snode, nothing, nothing, flat(ps.ϕ), flat(ps.γ))
    return y, st
end

# ---- MPPDEConv: m = phi([h_i; h_j; s_i - s_j; e_ij; theta]); y = psi([h_i; aggr(m); theta])   reference :390-422 ----
function (l::MPPDEConv)(x::CuF32Mat, ps, st::NamedTuple)
    g = st.graph
    snode = packed(g.ndata, Tuple(ksym(g.ndata)))
    edata = packed(g.edata, Tuple(ksym(g.edata)))
    theta = isempty(g.gdata) ? nothing : packed(map(v -> v isa AbstractVector ? reshape(v, :, 1) : v, g.gdata), Tuple(ksym(g.gdata)))
    theta === nothing || size(theta, 2) == g.num_graphs ||
        throw(DimensionMismatch("gdata has $(size(theta, 2)) columns but the batch holds $(g.num_graphs) graphs"))
    rows(a) = a === nothing ? 0 : size(a, 1)
    desc = ConvDesc(FAM_MPPDE, aggr_code(l.aggr), size(x, 1), rows(snode), 0, rows(edata), rows(theta), 0, 0, Mlp(l.ϕ), Mlp(l.ψ))
    y, _ = fused_conv(handle(g), desc, x, snode, edata, theta, flat(ps.ϕ), flat(ps.ψ))
    return y, st
end

# ---- GNOConv: m = reshape(phi([s_i; s_j; e_ij]), out, in) * h_j; y = act(W h_i + aggr(m) + b)   reference :509-547 ----
function (l::GNOConv)(x::CuF32Mat, ps, st::NamedTuple)
    g = st.graph
    snode = packed(g.ndata, Tuple(ksym(g.ndata)))
    edata = packed(g.edata, Tuple(ksym(g.edata)))
    rows(a) = a === nothing ? 0 : size(a, 1)
    desc = ConvDesc(FAM_GNO, aggr_code(l.aggr), size(x, 1), rows(snode), 0, rows(edata), 0, l.in_chs, l.out_chs, Mlp(l.ϕ), Mlp(l.linear))
    y, _ = fused_conv(handle(g), desc, x, snode, edata, nothing, flat(ps.ϕ), flat(ps.linear))
    return y, st
end

# ---- GCNConv: y = act(W (c .* A' (c .* x)) + b), c = 1/sqrt(in-degree)        reference src/layers.jl:200-239 ----
function (l::GCNConv)(x::CuF32Mat, ps, st::NamedTuple, edge_weight::Union{Nothing, CuVector{Float32}} = nothing)
    g = st.graph
    if edge_weight !== nothing
        length(edge_weight) == g.num_edges ||
            throw(AssertionError("Wrong number of edge weights (expected $(g.num_edges) but given $(length(edge_weight)))"))
    end
    size(x, 1) == l.in_chs || throw(DimensionMismatch("GCNConv($(l.in_chs) => $(l.out_chs)) got $(size(x, 1)) features"))
    gw = get_edge_weight(g)    # the graph's own stored weights (or nothing): they always weight the degree (layers.jl:224)
    desc = GcnDesc(l.in_chs, l.out_chs, act_code(l.activation), 1, l.add_self_loops, l.use_edge_weight)
    params = ps isa ComponentArray ? getdata(ps) : vcat(vec(ps.weight), vec(ps.bias))   # ps.bias is read unconditionally (:238)
    y = fused_gcn(handle(g), desc, x, params, edge_weight, gw)
    return y, st
end

# ---- SpectralConv: u'_i = 1/2 sum_j cos((x_i - x_j) n / 2) cot((x_i - x_j) / 2) u_j        reference src/layers.jl:652-662 ----
# The per-edge coefficient is static (a function of st.graph.edata.e and n only): it is evaluated once per graph in the precision
# `e` is stored in, rounded to Float32 and cached; the call is one pass of the ordered aggregate kernel.  The pullback w.r.t. u is
# the same kernel on the transposed graph (edges kept in their stored order, so the coefficient vector is shared).
const SPECTRAL_COEF = IdDict{Any, CuVector{Float32}}()
function spectral_coef(l::SpectralConv, g::GNNGraph)
    get!(SPECTRAL_COEF, g.edata.e) do
        e = Float64.(vec(Array(g.edata.e)))
        CuVector{Float32}(Float32.(cos.(e .* l.n ./ 2) .* cot.(e ./ 2) ./ 2))
    end
end

function weighted_sum(h::GraphHandle, ht::GraphHandle, x::CuF32Mat, w::CuVector{Float32})
    y = similar(x)
    aggregate!(h.ptr, AGGR_SUM, x, Int32(size(x, 1)), w, y)
    return y
end

function ChainRulesCore.rrule(::typeof(weighted_sum), h::GraphHandle, ht::GraphHandle, x::CuF32Mat, w::CuVector{Float32})
    y = weighted_sum(h, ht, x, w)
    function weighted_sum_pullback(dy)
        dx = similar(x)
        aggregate!(ht.ptr, AGGR_SUM, CuMatrix{Float32}(unthunk(dy)), Int32(size(x, 1)), w, dx)
        return NoTangent(), NoTangent(), NoTangent(), dx, NoTangent()
    end
    return y, weighted_sum_pullback
end

function (l::SpectralConv)(x::CuF32Mat, ps, st::NamedTuple)
    g = st.graph
    s, t = edge_index(g)
    gt = GNNGraph(t, s; num_nodes = g.num_nodes)          # reversed edges, same stored order (its handle is cached like any other)
    return weighted_sum(handle(g), handle(gt), x, spectral_coef(l, g)), st
end

function (l::SpectralConv)(x::CuVector{Float32}, ps, st::NamedTuple)
    y, st = l(reshape(x, 1, :), ps, st)
    return vec(y), st
end
