"""GNNGraph mirror: the COO graph the reference keeps in `st.graph`, plus the cached device layout behind it.

Mirrors the parts of GraphNeuralNetworks.GNNGraph the reference touches ([DEP]):
`GNNGraph(s, t)`, `GNNGraph(g; ndata, edata, gdata)`, `num_nodes/num_edges/num_graphs`, `ndata/edata/gdata`,
`batch`, `add_self_loops`, `rand_graph`, and `copy` (/root/reference/src/utils.jl:8).

Arrays use Julia shapes: node data `(D, N)`, edge data `(D, E)`, graph data `(D, G)`; a feature matrix is stored
column-major, i.e. `x.T` is a contiguous `[N, D]` torch tensor (`from_rowmajor` / `rowmajor` convert for free).
Indices are 0-based on this side of the boundary (`index_base=1` accepts Julia's).

The CSR layout (libngpde graph handle) is built once per topology, lazily, and shared by every shallow copy of the
graph -- it is NOT part of `st`, so `st == (graph=g,)` keeps holding after a call (test/runtests.jl:21-24).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib

Tensor = torch.Tensor


def from_rowmajor(t: Tensor) -> Tensor:
    """[N, D] row-major buffer -> Julia-shaped (D, N) column-major view (no copy)."""
    return t.T


def rowmajor(x: Tensor, dtype=torch.float32) -> Tensor:
    """Julia-shaped (D, N) -> contiguous [N, D] float32 (no copy when x is already column-major float32)."""
    if x.dim() == 1:
        x = x.reshape(1, -1)
    r = x.T
    if r.dtype != dtype:
        r = r.to(dtype)
    return r.contiguous()


class GraphHandle:
    """Owner of one native `ngpde_graph_t` (the device CSR layout of one topology on one device).  It is what the layers
    pass to the C ABI (`_as_parameter_` makes ctypes take the raw pointer) and what autograd contexts and runners keep a
    reference to: the native layout lives exactly as long as somebody can still call into it -- a graph that is garbage
    collected, or asked for a layout on another device, between a forward and its backward cannot free the arrays the
    backward kernels read."""

    def __init__(self, ptr: C.c_void_p, device: torch.device):
        self._as_parameter_ = ptr
        self.device = device

    @property
    def value(self):
        return self._as_parameter_.value

    def __del__(self):
        ptr, self._as_parameter_ = getattr(self, "_as_parameter_", None), None
        if ptr is not None and ptr.value:
            try:
                _lib.load().ngpde_graph_destroy(ptr)
            except Exception:
                pass


class _Topology:
    """Edge lists + the lazily built libngpde handles (one per device); shared between shallow copies of a graph."""

    def __init__(self, s: Tensor, t: Tensor, num_nodes: int, num_graphs: int):
        self.s, self.t = s, t
        self.num_nodes, self.num_graphs = int(num_nodes), int(num_graphs)
        self._handles: Dict[torch.device, GraphHandle] = {}

    def handle(self, device: torch.device) -> GraphHandle:
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        hit = self._handles.get(device)
        if hit is not None:
            return hit
        if device.type != "cuda":
            raise _lib.NgpdeError("the message-passing path runs on CUDA only (no CPU fallback); move x/ps/st to a GPU")
        lib = _lib.load()
        s = self.s.to(device=device, dtype=torch.int64).contiguous()
        t = self.t.to(device=device, dtype=torch.int64).contiguous()
        h = C.c_void_p()
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(lib.ngpde_graph_create(C.byref(h), self.num_nodes, s.numel(), s.data_ptr(), t.data_ptr(),
                                              _lib.IDX_I64, 0, 1, self.num_graphs, stream))
        self._handles[device] = GraphHandle(h, device)
        return self._handles[device]

    def array(self, name: str, device: torch.device, with_self_loops: bool = False) -> Tensor:
        """Copy of one of the handle's integer arrays (bit-exact index contract; used by tests)."""
        lib = _lib.load()
        h = self.handle(device)
        ptr, n = C.c_void_p(), C.c_int64()
        with torch.cuda.device(device):
            _lib.check(lib.ngpde_graph_array(h, _lib.GA[name], int(with_self_loops), C.byref(ptr), C.byref(n)))
            out = torch.empty(n.value, dtype=torch.int32, device=device)
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(lib.ngpde_graph_array_copy(h, _lib.GA[name], int(with_self_loops), out.data_ptr(), n.value,
                                                  stream))
        return out


def _as_index(a, base: int) -> Tensor:
    if isinstance(a, Tensor):
        t = a.to(torch.int64)
    else:
        t = torch.as_tensor(np.asarray(a), dtype=torch.int64)
    return t - base if base else t


def _norm_data(d, default_key: str) -> Dict[str, Tensor]:
    if d is None:
        return {}
    if isinstance(d, Tensor):
        return {default_key: d}  # a bare array is stored under :x / :e ([DEP]; test/runtests.jl:145,197-199)
    return dict(d)


class GNNGraph:
    def __init__(self, s: Union["GNNGraph", Tensor, Sequence[int]], t=None, *, num_nodes: Optional[int] = None,
                 ndata=None, edata=None, gdata=None, num_graphs: int = 1, w: Optional[Tensor] = None,
                 index_base: int = 0):
        if isinstance(s, GNNGraph):  # GNNGraph(g; ndata=..., edata=..., gdata=...): shallow copy, arrays shared
            g = s
            self._topo = g._topo
            self.ndata = _norm_data(ndata, "x") if ndata is not None else g.ndata
            self.edata = _norm_data(edata, "e") if edata is not None else g.edata
            self.gdata = _norm_data(gdata, "u") if gdata is not None else g.gdata
            self.w = g.w if w is None else w
        else:
            si, ti = _as_index(s, index_base), _as_index(t, index_base)
            if si.shape != ti.shape or si.dim() != 1:
                raise ValueError("s and t must be 1-D index vectors of equal length")
            if num_nodes is None:
                num_nodes = int(max(si.max().item(), ti.max().item())) + 1 if si.numel() else 0
            self._topo = _Topology(si, ti, num_nodes, num_graphs)
            self.ndata, self.edata, self.gdata = _norm_data(ndata, "x"), _norm_data(edata, "e"), _norm_data(gdata, "u")
            self.w = w
        self._cache: Dict = {}

    def transposed(self) -> "GNNGraph":
        """The graph with every edge reversed, edges in the same stored order (cached; topology only)."""
        gt = self._cache.get("transposed")
        if gt is None:
            gt = GNNGraph(self.t, self.s, num_nodes=self.num_nodes, num_graphs=self.num_graphs)
            self._cache["transposed"] = gt
        return gt

    # --- GNNGraph fields ---
    @property
    def s(self) -> Tensor:
        return self._topo.s

    @property
    def t(self) -> Tensor:
        return self._topo.t

    @property
    def num_nodes(self) -> int:
        return self._topo.num_nodes

    @property
    def num_edges(self) -> int:
        return int(self._topo.s.numel())

    @property
    def num_graphs(self) -> int:
        return self._topo.num_graphs

    def edge_index(self):
        return self.s, self.t

    def __eq__(self, other) -> bool:
        if not isinstance(other, GNNGraph):
            return NotImplemented
        if self.num_nodes != other.num_nodes or self.num_graphs != other.num_graphs:
            return False
        if self._topo is not other._topo:
            if self.s.shape != other.s.shape or not (torch.equal(self.s.cpu(), other.s.cpu()) and
                                                     torch.equal(self.t.cpu(), other.t.cpu())):
                return False
        for a, b in ((self.ndata, other.ndata), (self.edata, other.edata), (self.gdata, other.gdata)):
            if list(a.keys()) != list(b.keys()):
                return False
            for k in a:
                if a[k] is not b[k] and not (a[k].shape == b[k].shape and torch.equal(a[k].cpu(), b[k].cpu())):
                    return False
        return True

    __hash__ = None

    def __repr__(self) -> str:
        return (f"GNNGraph(num_nodes={self.num_nodes}, num_edges={self.num_edges}, num_graphs={self.num_graphs}, "
                f"ndata={list(self.ndata)}, edata={list(self.edata)}, gdata={list(self.gdata)})")

    def to(self, device) -> "GNNGraph":
        """`g |> gpu`: move edge lists and data arrays."""
        device = torch.device(device)
        g = GNNGraph.__new__(GNNGraph)
        if self._topo.s.device == device:
            g._topo = self._topo
        else:
            g._topo = _Topology(self.s.to(device), self.t.to(device), self.num_nodes, self.num_graphs)
        mv = lambda d: {k: v.to(device) for k, v in d.items()}
        g.ndata, g.edata, g.gdata = mv(self.ndata), mv(self.edata), mv(self.gdata)
        g.w = None if self.w is None else self.w.to(device)
        g._cache = {}
        return g

    # --- device-side views used by the layers ---
    def handle(self, device: torch.device):
        return self._topo.handle(device)

    def layout_array(self, name: str, device=None, with_self_loops: bool = False) -> Tensor:
        device = torch.device(device) if device is not None else self.s.device
        return self._topo.array(name, device, with_self_loops)

    def _packed(self, kind: str, keys: Sequence[str], source: Dict[str, Tensor], device, width_axis0: bool = True):
        """[items, sum D] float32 row-major concatenation of the named fields.  Cached per `kind`; the entry keeps the
        SOURCE tensors alive and is valid only while the very same tensor objects, at the same in-place version, are
        still installed (identity, not address: a freed block handed out again by the caching allocator cannot alias)."""
        srcs = tuple(source[k] for k in keys)
        vers = tuple(t._version for t in srcs)
        hit = self._cache.get(kind)
        if (hit is not None and hit[0] == (tuple(keys), str(device)) and len(hit[1]) == len(srcs)
                and all(a is b for a, b in zip(hit[1], srcs)) and hit[2] == vers):
            return hit[3]
        if not keys:
            packed = None
        else:
            for k, t in zip(keys, srcs):
                _warn_float64(k, t)
            parts = [rowmajor(t.to(device)) for t in srcs]
            packed = parts[0] if len(parts) == 1 else torch.cat(parts, dim=1).contiguous()
        self._cache[kind] = ((tuple(keys), str(device)), srcs, vers, packed)
        return packed


_F64_WARNED = set()


def _warn_float64(name: str, t: Tensor) -> None:
    """Float64 side data (the reference's tests build `rand(2, n)` arrays: test/runtests.jl:58-61,126-128).  Julia promotes
    the whole layer call to Float64 there; this path computes in float32 only, so the data is converted -- once per
    field name, loudly."""
    if t.dtype == torch.float64 and name not in _F64_WARNED:
        import warnings
        _F64_WARNED.add(name)
        warnings.warn(f"graph data field {name!r} is float64: converted to float32 (the B200 path computes in float32; the "
                      "reference would promote the layer call to Float64)", stacklevel=3)


def copy(g: GNNGraph, **kwargs) -> GNNGraph:
    """Base.copy(g::GNNGraph; kwargs...) = GNNGraph(g; kwargs...)  (/root/reference/src/utils.jl:8)."""
    return GNNGraph(g, **kwargs)


def add_self_loops(g: GNNGraph) -> GNNGraph:
    """[DEP] appends (i, i) for every node at the END of the COO lists."""
    n = g.num_nodes
    loops = torch.arange(n, dtype=torch.int64, device=g.s.device)
    w = None if g.w is None else torch.cat([g.w, torch.ones(n, dtype=g.w.dtype, device=g.w.device)])
    return GNNGraph(torch.cat([g.s, loops]), torch.cat([g.t, loops]), num_nodes=n, ndata=g.ndata, gdata=g.gdata,
                    num_graphs=g.num_graphs, w=w)


def batch(gs: Sequence[GNNGraph]) -> GNNGraph:
    """[DEP] MLUtils.batch: block-diagonal concatenation; gdata concatenated along the graph axis."""
    s, t, off = [], [], 0
    for g in gs:
        s.append(g.s + off)
        t.append(g.t + off)
        off += g.num_nodes
    cat = lambda name: {k: torch.cat([getattr(g, name)[k] for g in gs], dim=1) for k in getattr(gs[0], name)}
    gd = {}
    for k in gs[0].gdata:
        gd[k] = torch.cat([g.gdata[k] if g.gdata[k].dim() == 2 else g.gdata[k].reshape(-1, 1) for g in gs], dim=1)
    return GNNGraph(torch.cat(s), torch.cat(t), num_nodes=off, ndata=cat("ndata"), edata=cat("edata"), gdata=gd,
                    num_graphs=sum(g.num_graphs for g in gs))


def rand_graph(n: int, m: int, *, bidirected: bool = True, seed: Optional[int] = None) -> GNNGraph:
    """[DEP] GNNGraphs.rand_graph(n, m): m random edges (m/2 pairs emitted both ways when bidirected)."""
    rng = np.random.default_rng(seed)
    if n == 0 or m == 0:
        z = torch.zeros(0, dtype=torch.int64)
        return GNNGraph(z, z.clone(), num_nodes=n)
    if bidirected:
        half = m // 2
        a, b = rng.integers(0, n, half), rng.integers(0, n, half)
        s, t = np.concatenate([a, b]), np.concatenate([b, a])
    else:
        s, t = rng.integers(0, n, m), rng.integers(0, n, m)
    return GNNGraph(torch.from_numpy(s), torch.from_numpy(t), num_nodes=n)
