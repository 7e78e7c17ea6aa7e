# The two differentiable primitives every layer method reduces to, and their ChainRules `rrule`s.  Zygote sees one opaque
# call per layer; the pullback is the matching hand-written backward kernel (ngpde_*_backward).

workspace(nbytes::Integer) = CuArray{UInt8}(undef, max(Int(nbytes), 256))
dm(desc::ConvDesc) = desc.family == FAM_GNO ? Int(desc.gno_out) : Int(desc.phi.dims[desc.phi.n_layers + 1])   # rows of mbar
dy(desc::ConvDesc) = Int(desc.node.dims[desc.node.n_layers + 1])                                             # rows of y

"""
    fused_conv(h, desc, x, snode, edata, theta, phi, node) -> (y, mbar)

One message-passing layer call (`ngpde_conv_forward`): `x (dx, N)`, static node data `snode (dhs+dpos, N)`, edge data
`edata (de, E)` in stored edge order, `theta (dθ, G)`, flat parameter segments `phi`, `node`.
"""
function fused_conv(h::GraphHandle, desc::ConvDesc, x::CuMatrix{Float32}, snode, edata, theta, phi::CuVector{Float32}, node)
    N = size(x, 2)
    has_node = desc.node.n_layers > 0
    mbar = CuMatrix{Float32}(undef, dm(desc), N)
    y = has_node ? CuMatrix{Float32}(undef, dy(desc), N) : mbar
    nb = conv_workspace_bytes(h, desc, false)
    nb == 0 && check(-1)
    ws = workspace(nb)
    io = ConvIO(pointer(x), ptr(snode), ptr(edata), ptr(theta), pointer(phi), ptr(node), pointer(mbar), pointer(y),
                CU_NULL, CU_NULL, CU_NULL, CU_NULL, CU_NULL)
    GC.@preserve h x snode edata theta phi node mbar y ws conv_forward!(h.ptr, desc, io, ws)
    return y, mbar
end

function ChainRulesCore.rrule(::typeof(fused_conv), h::GraphHandle, desc::ConvDesc, x, snode, edata, theta, phi, node)
    y, mbar = fused_conv(h, desc, x, snode, edata, theta, phi, node)
    has_node = desc.node.n_layers > 0
    function fused_conv_pullback(Δ)
        dy = unthunk(Δ[1])     # the cotangent of mbar (second output) is never used by the layer methods
        dy isa AbstractZero && return ntuple(_ -> NoTangent(), 9)
        dyv = dy isa CuMatrix{Float32} ? dy : CuMatrix{Float32}(dy)
        dx, dphi = similar(x), similar(phi)
        dnode = has_node ? similar(node) : nothing
        nb = conv_workspace_bytes(h, desc, true)
        nb == 0 && check(-1)
        ws = workspace(nb)
        io = ConvIO(pointer(x), ptr(snode), ptr(edata), ptr(theta), pointer(phi), ptr(node), pointer(mbar), pointer(y),
                    pointer(dyv), pointer(dx), pointer(dphi), ptr(dnode), CU_NULL)
        GC.@preserve h x snode edata theta phi node mbar y dyv dx dphi dnode ws conv_backward!(h.ptr, desc, io, ws)
        # tangents for (x, phi, node); the graph, the static data and theta carry none (the reference wraps theta in
        # `@ignore_derivatives`, src/layers.jl:397, and `ndata` never enters `ps`)
        return NoTangent(), NoTangent(), NoTangent(), dx, NoTangent(), NoTangent(), NoTangent(), dphi,
               (has_node ? dnode : NoTangent())
    end
    return (y, mbar), fused_conv_pullback
end

"""
    fused_gcn(h, desc, x, params, edge_weight, graph_weight) -> y

GCNConv (`ngpde_gcn_conv_forward`); `params = [vec(weight); vec(bias)]` flat, exactly `ComponentArray(ps)`.
"""
function fused_gcn(h::GraphHandle, desc::GcnDesc, x::CuMatrix{Float32}, params::CuVector{Float32}, ew, gw)
    N = size(x, 2)
    y = CuMatrix{Float32}(undef, Int(desc.out_chs), N)
    nb = gcn_workspace_bytes(h, desc, false)
    nb == 0 && check(-1)
    ws = workspace(nb)
    nw = Int(desc.in_chs) * Int(desc.out_chs)
    GC.@preserve h x params ew gw y ws gcn_forward!(h.ptr, desc, pointer(x), pointer(params), pointer(params, nw + 1), ptr(ew), ptr(gw),
                                                    pointer(y), ws)
    return y
end

function ChainRulesCore.rrule(::typeof(fused_gcn), h::GraphHandle, desc::GcnDesc, x, params, ew, gw)
    y = fused_gcn(h, desc, x, params, ew, gw)
    function fused_gcn_pullback(Δ)
        dy = unthunk(Δ)
        dyv = dy isa CuMatrix{Float32} ? dy : CuMatrix{Float32}(dy)
        dx, dparams = similar(x), CUDA.zeros(Float32, length(params))
        nb = gcn_workspace_bytes(h, desc, true)
        nb == 0 && check(-1)
        ws = workspace(nb)
        nw = Int(desc.in_chs) * Int(desc.out_chs)
        GC.@preserve h x params ew gw y dyv dx dparams ws gcn_backward!(h.ptr, desc, pointer(x), pointer(params),
            pointer(params, nw + 1), ptr(ew), ptr(gw), pointer(y), pointer(dyv), pointer(dx), pointer(dparams), pointer(dparams, nw + 1), ws)
        # edge weights are not differentiated (GNN.jl treats them as data in GCNConv's normalisation as well)
        return NoTangent(), NoTangent(), NoTangent(), dx, dparams, NoTangent(), NoTangent()
    end
    return y, fused_gcn_pullback
end

# flat Float32 device vector of a parameter (sub)tree in field order: zero-copy for a ComponentArray view (the tutorials'
# `ps = ComponentArray(ps) |> gpu`, docs/src/tutorials/graph_node.md:90), a differentiable `vcat(vec.(leaves)...)` otherwise
flat(ps::ComponentArray) = getdata(ps)
flat(ps::CuVector{Float32}) = ps
flat(ps::NamedTuple) = reduce(vcat, map(flatleaf, leaves(ps)))
flatleaf(a::AbstractArray) = vec(a)
leaves(nt::NamedTuple) = reduce((acc, v) -> v isa NamedTuple ? (acc..., leaves(v)...) : (acc..., v), values(nt); init = ())
