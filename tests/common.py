"""Shared helpers for the parity tests: product (ngpde) <-> oracle conversions and error measures."""
import numpy as np
import torch

import ngpde
import ngpde_oracle as orc


# the product's NT stores field names NFKC-normalised (U+03C6); the oracle spells them as the reference does (U+03D5)
_ORACLE_KEY = {"\u03c6": "\u03d5"}


def relerr(a: torch.Tensor, b: torch.Tensor) -> float:
    """Max-norm relative error  max|a-b| / max|b|  (the tolerance measure of every float parity test)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.numel() == 0:
        return 0.0
    den = max(b.abs().max().item(), 1e-30)
    return (a - b).abs().max().item() / den


def block_relerrs(a: torch.Tensor, b: torch.Tensor, blocks) -> dict:
    """Max-norm relative error per named block: blocks = [(name, index-or-slice into the leading axis / flat vector)].
    A block whose reference is identically zero must be matched to 1e-30 absolute (den = max|b| of the WHOLE array
    there, so that e.g. the zero gradient of an unused bias does not divide by zero)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    whole = max(b.abs().max().item(), 1e-30) if b.numel() else 1.0
    out = {}
    for name, sel in blocks:
        aa, bb = a[sel], b[sel]
        if bb.numel() == 0:
            continue
        den = bb.abs().max().item()
        if den == 0.0:
            den = whole
        out[name] = (aa - bb).abs().max().item() / den
    return out


def param_blocks(tree, prefix=""):
    """[(name, slice)] of the leaves of a Julia-shaped parameter tree in ComponentArray (flat) order."""
    out, off = [], 0

    def walk(t, pre):
        nonlocal off
        for k, v in t.items():
            if isinstance(v, torch.Tensor):
                out.append((pre + k, slice(off, off + v.numel())))
                off += v.numel()
            else:
                walk(v, pre + k + ".")

    walk(tree, prefix)
    return out


def tree_to_cpu(ps, dtype=torch.float32):
    """ngpde NT / ComponentArray tree -> nested dict of CPU tensors (Julia shapes) for the oracle."""
    if isinstance(ps, ngpde.ComponentArray):
        ps = ps.to_tree()
    out = {}
    for k, v in ps.items():
        k = _ORACLE_KEY.get(k, k)
        out[k] = v.detach().cpu().to(dtype).contiguous().clone() if isinstance(v, torch.Tensor) else tree_to_cpu(v, dtype)
    return out


def tree_requires_grad(tree):
    for v in tree.values():
        if isinstance(v, torch.Tensor):
            v.requires_grad_(True)
        else:
            tree_requires_grad(v)
    return tree


def tree_leaves(tree):
    for v in tree.values():
        if isinstance(v, torch.Tensor):
            yield v
        else:
            yield from tree_leaves(v)


def flat_grad(tree) -> torch.Tensor:
    """Gradients of a Julia-shaped parameter tree flattened in ComponentArray order (column-major leaves)."""
    parts = []
    for v in tree_leaves(tree):
        g = v.grad if v.grad is not None else torch.zeros_like(v)
        parts.append(g.T.reshape(-1) if g.dim() == 2 else g.reshape(-1))
    return torch.cat(parts)


def to_ograph(g: ngpde.GNNGraph, dtype=torch.float32) -> orc.OGraph:
    cv = lambda d: {k: v.detach().cpu().to(dtype) for k, v in d.items()}
    return orc.OGraph(g.s.cpu().numpy().astype(np.int64), g.t.cpu().numpy().astype(np.int64), g.num_nodes, g.num_graphs,
                      cv(g.ndata), cv(g.edata), cv(g.gdata), None if g.w is None else g.w.detach().cpu().to(dtype))


def jl_rand(rng: np.random.Generator, d: int, n: int, device="cpu", lo=-1.0, hi=1.0) -> torch.Tensor:
    """(d, n) Julia-shaped float32 tensor stored column-major, U(lo, hi)."""
    a = rng.uniform(lo, hi, size=(n, d)).astype(np.float32)
    return torch.from_numpy(a).to(device).T


def random_graph(rng, n, e, device="cpu", **kw):
    s = torch.from_numpy(rng.integers(0, n, e))
    t = torch.from_numpy(rng.integers(0, n, e))
    return ngpde.GNNGraph(s.to(device), t.to(device), num_nodes=n, **kw)


def oracle_forward(layer, x: torch.Tensor, ps: dict, og: orc.OGraph, edge_weight=None) -> torch.Tensor:
    """Evaluate the oracle's restatement of `layer` (an ngpde layer object or a Chain of them) on CPU tensors.
    `ps` is the tree produced by tree_to_cpu (oracle key spelling)."""
    from ngpde.lux import mlp_spec
    if isinstance(layer, ngpde.Chain):
        for i, l in enumerate(layer.layers):
            x = oracle_forward(l, x, ps[f"layer_{i + 1}"], og)
        return x
    if isinstance(layer, ngpde.ExplicitEdgeConv):
        return orc.explicit_edge_conv(x, ps, og, mlp_spec(layer.ϕ), layer.aggr)
    if isinstance(layer, ngpde.VMHConv):
        return orc.vmh_conv(x, ps, og, mlp_spec(layer.ϕ), mlp_spec(layer.γ), layer.aggr)
    if isinstance(layer, ngpde.MPPDEConv):
        return orc.mppde_conv(x, ps, og, mlp_spec(layer.ϕ), mlp_spec(layer.ψ), layer.aggr)
    if isinstance(layer, ngpde.GNOConv):
        return orc.gno_conv(x, ps, og, layer.in_chs, layer.out_chs, mlp_spec(layer.ϕ), layer.linear.activation,
                            layer.aggr, layer.bias)
    if isinstance(layer, ngpde.GCNConv):
        return orc.gcn_conv(x, ps, og, layer.in_chs, layer.out_chs, layer.activation, layer.add_self_loops,
                            layer.use_edge_weight, edge_weight, layer.bias)
    raise TypeError(type(layer))


def product_fwd_bwd(layer, x, ps, st, dy=None, **kw):
    """y, dx, flat dps of the CUDA path (ps: NT tree on the device; gradients in ComponentArray order)."""
    x = x.detach().clone().requires_grad_(True)
    ca = ngpde.ComponentArray(ps)
    ca.data.requires_grad_(True)
    y, _ = layer(x, ca, st, **kw)
    if dy is None:
        return y.detach(), None, None
    y.backward(dy)
    return y.detach(), x.grad.detach(), ca.data.grad.detach()


def oracle_fwd_bwd(layer, x, ps, g, dy=None, dtype=torch.float32, **kw):
    """Same through the oracle (torch autograd over the unfused restatement), on CPU in `dtype`."""
    og = to_ograph(g, dtype)
    xc = x.detach().cpu().to(dtype).clone().requires_grad_(dy is not None)
    pc = tree_to_cpu(ps, dtype)
    if dy is not None:
        tree_requires_grad(pc)
    y = oracle_forward(layer, xc, pc, og, **kw)
    if dy is None:
        return y.detach(), None, None
    y.backward(dy.detach().cpu().to(y.dtype))
    return y.detach(), xc.grad.detach(), flat_grad(pc)
