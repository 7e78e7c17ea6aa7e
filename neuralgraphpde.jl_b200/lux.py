"""Minimal mirror of the Lux explicit-layer protocol the reference's layers live in ([DEP] Lux 0.4):
`Dense`, `Chain`, `setup(rng, layer) -> (ps, st)`, NamedTuple-like parameter/state trees and `ComponentArray`.

Parameters keep Julia's shapes and memory order: `weight` is `(out, in)` column-major (so `weight.T` is the
contiguous `[in][out]` buffer libngpde consumes), `bias` is `(out, 1)`.  `ComponentArray(ps)` is the flat vector in
NamedTuple field order (`graph_node.md:90`, `VMH.md:127`); sub-trees of it are zero-copy views, so a layer called
with `ps::ComponentArray` hands its flat segment to the CUDA library without packing.
"""
from __future__ import annotations

import math
import unicodedata
from typing import Any, Callable, Dict, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

Tensor = torch.Tensor
MlpSpec = List[Tuple[int, int, str, bool]]


def nfkc(name: str) -> str:
    """Python NFKC-normalises identifiers, so `ps.ϕ` (U+03D5, the reference's field name) arrives as U+03C6."""
    return unicodedata.normalize("NFKC", name)


def _lookup(mapping, k):
    if k in mapping:
        return k
    nk = nfkc(k)
    for key in mapping:
        if nfkc(key) == nk:
            return key
    return None


class NT(dict):
    """Ordered NamedTuple stand-in: `nt.field`, `nt == other`, iteration over values like Julia's `values(nt)`.

    Field names are stored NFKC-normalised (what Python does to identifiers anyway), so the reference's `ϕ`
    (U+03D5) is kept as `φ` (U+03C6); lookups accept either spelling."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            dict.__setitem__(self, nfkc(k), v)

    def __setitem__(self, k, v):
        raise TypeError("NamedTuples are immutable; use merge(nt, NT(k=v))")

    def __getattr__(self, k):
        key = _lookup(self, k)
        if key is None:
            raise AttributeError(k)
        return dict.__getitem__(self, key)

    def __getitem__(self, k):
        key = _lookup(self, k)
        if key is None:
            raise KeyError(k)
        return dict.__getitem__(self, key)

    def __setattr__(self, k, v):
        raise TypeError("NamedTuples are immutable; use merge(nt, NT(k=v))")

    def __eq__(self, other):
        if not isinstance(other, dict) or list(self.keys()) != list(other.keys()):
            return False
        for k in self:
            a, b = self[k], other[k]
            if isinstance(a, Tensor) or isinstance(b, Tensor):
                if not (isinstance(a, Tensor) and isinstance(b, Tensor) and a.shape == b.shape and torch.equal(a, b)):
                    return False
            elif not (a == b):
                return False
        return True

    def __ne__(self, other):
        return not self.__eq__(other)

    __hash__ = None

    def __repr__(self):
        return "(" + ", ".join(f"{k} = {v!r}" if not isinstance(v, Tensor) else f"{k} = Tensor{tuple(v.shape)}"
                               for k, v in self.items()) + ("," if len(self) == 1 else "") + ")"


def merge(a: NT, b: dict) -> NT:
    out = NT(a)
    for k, v in b.items():
        dict.__setitem__(out, nfkc(k), v)
    return out


# ---- initialisers (Lux.glorot_uniform / glorot_normal / zeros32 / ones32) ----

def _np_rng(rng) -> np.random.Generator:
    if isinstance(rng, np.random.Generator):
        return rng
    return np.random.default_rng(rng)


def glorot_uniform(rng, out_dims: int, in_dims: int) -> np.ndarray:
    a = math.sqrt(6.0 / (in_dims + out_dims))
    return _np_rng(rng).uniform(-a, a, size=(out_dims, in_dims)).astype(np.float32)


def glorot_normal(rng, out_dims: int, in_dims: int) -> np.ndarray:
    std = math.sqrt(2.0 / (in_dims + out_dims))
    return (_np_rng(rng).standard_normal(size=(out_dims, in_dims)) * std).astype(np.float32)


def zeros32(rng, *dims) -> np.ndarray:
    return np.zeros(dims, dtype=np.float32)


def ones32(rng, *dims) -> np.ndarray:
    return np.ones(dims, dtype=np.float32)


def julia_array(a: np.ndarray, device="cpu") -> Tensor:
    """numpy (d0, d1) -> torch tensor of the same shape stored column-major (Julia memory order)."""
    a = np.asarray(a, dtype=np.float32)
    if a.ndim == 1:
        return torch.from_numpy(a.copy()).to(device)
    return torch.from_numpy(np.ascontiguousarray(a.T)).to(device).T


ACT_NAMES = {"identity", "relu", "tanh", "sigmoid", "swish", "gelu", "softplus", "elu", "leakyrelu"}


def _act_name(act) -> str:
    """Accept NNlib-style names or torch callables; `tanh_fast`/`sigmoid_fast` (NNlib.fast_act) map to the exact ones."""
    if act is None:
        return "identity"
    if isinstance(act, str):
        name = {"tanh_fast": "tanh", "sigmoid_fast": "sigmoid", "σ": "sigmoid", "silu": "swish"}.get(act, act)
    else:
        name = {torch.tanh: "tanh", torch.relu: "relu", torch.sigmoid: "sigmoid",
                torch.nn.functional.relu: "relu", torch.nn.functional.silu: "swish",
                torch.nn.functional.gelu: "gelu", torch.nn.functional.softplus: "softplus",
                torch.nn.functional.elu: "elu", torch.nn.functional.leaky_relu: "leakyrelu"}.get(act)
        if name is None:
            raise ValueError(f"unsupported activation {act!r}")
    if name not in ACT_NAMES:
        raise ValueError(f"unsupported activation {name!r}; supported: {sorted(ACT_NAMES)}")
    return name


class AbstractExplicitLayer:
    def initialparameters(self, rng, device="cpu") -> NT:
        return NT()

    def initialstates(self, rng) -> NT:
        return NT()

    def parameterlength(self) -> int:
        return 0

    def statelength(self) -> int:
        return 0


class Dense(AbstractExplicitLayer):
    """Lux.Dense(in => out, activation; bias=true, init_weight=glorot_uniform, init_bias=zeros32)."""

    def __init__(self, in_dims: Union[int, Tuple[int, int]], out_dims: Optional[int] = None, activation="identity", *,
                 bias: bool = True, init_weight: Callable = glorot_uniform, init_bias: Callable = zeros32):
        if isinstance(in_dims, tuple):  # Dense((4, 5), act) stands for Julia's `Dense(4 => 5, act)`
            if out_dims is not None:
                activation = out_dims
            in_dims, out_dims = in_dims
        self.in_dims, self.out_dims = int(in_dims), int(out_dims)
        self.activation = _act_name(activation)
        self.bias = bool(bias)
        self.init_weight, self.init_bias = init_weight, init_bias

    def initialparameters(self, rng, device="cpu") -> NT:
        p = NT(weight=julia_array(self.init_weight(rng, self.out_dims, self.in_dims), device))
        if self.bias:
            dict.__setitem__(p, "bias", julia_array(self.init_bias(rng, self.out_dims, 1), device))
        return p

    def parameterlength(self) -> int:
        return self.out_dims * (self.in_dims + (1 if self.bias else 0))

    def spec(self) -> MlpSpec:
        return [(self.in_dims, self.out_dims, self.activation, self.bias)]

    def __repr__(self):
        a = "" if self.activation == "identity" else f", {self.activation}"
        return f"Dense({self.in_dims} => {self.out_dims}{a})"


class Chain(AbstractExplicitLayer):
    """Lux.Chain(layers...): ps = (layer_1 = ..., layer_2 = ...), st likewise."""

    def __init__(self, *layers):
        self.layers = list(layers)

    def initialparameters(self, rng, device="cpu") -> NT:
        return NT((f"layer_{i + 1}", l.initialparameters(rng, device)) for i, l in enumerate(self.layers))

    def initialstates(self, rng) -> NT:
        return NT((f"layer_{i + 1}", l.initialstates(rng)) for i, l in enumerate(self.layers))

    def parameterlength(self) -> int:
        return sum(l.parameterlength() for l in self.layers)

    def is_mlp(self) -> bool:
        return all(isinstance(l, Dense) for l in self.layers)

    def spec(self) -> MlpSpec:
        if not self.is_mlp():
            raise TypeError("only Chains of Dense layers can be fused into the message-passing kernels")
        out: MlpSpec = []
        for l in self.layers:
            out += l.spec()
        return out

    def __call__(self, x, ps, st):
        """Generic Chain call (used for Chains of graph layers, e.g. GCNConv -> GCNConv)."""
        new_st = NT()
        for i, l in enumerate(self.layers):
            k = f"layer_{i + 1}"
            x, s = l(x, getattr(ps, k), st[k])
            dict.__setitem__(new_st, k, s)
        return x, new_st

    def __repr__(self):
        return "Chain(" + ", ".join(map(repr, self.layers)) + ")"


def mlp_spec(layer) -> MlpSpec:
    if isinstance(layer, (Dense, Chain)):
        return layer.spec()
    raise TypeError(f"expected a Dense or a Chain of Dense layers, got {type(layer).__name__}")


def setup(rng, layer, device="cpu") -> Tuple[NT, NT]:
    """Lux.setup(rng, layer) -> (ps, st).  `device="cuda"` plays the role of `|> gpu`."""
    rng = _np_rng(rng)
    return layer.initialparameters(rng, device), layer.initialstates(rng)


# ---- ComponentArray ----

def _leaves(tree, prefix=()) -> Iterator[Tuple[Tuple[str, ...], Tensor]]:
    for k, v in tree.items():
        if isinstance(v, Tensor):
            yield prefix + (k,), v
        else:
            yield from _leaves(v, prefix + (k,))


def _flat_leaf(v: Tensor) -> Tensor:
    """Column-major flattening of a Julia-shaped array (vec(A) in Julia)."""
    return v.T.reshape(-1) if v.dim() == 2 else v.reshape(-1)


class ComponentArray:
    """Flat parameter vector with named, zero-copy sub-views (ComponentArrays.jl as used in graph_node.md:90)."""

    def __init__(self, tree_or_data, axes: Optional[Dict] = None):
        if axes is None:
            tree = tree_or_data
            parts: List[Tensor] = []
            object.__setattr__(self, "axes", self._build_axes(tree, parts))
            data = torch.cat(parts) if parts else torch.zeros(0)
            object.__setattr__(self, "data", data.detach().clone())
        else:
            object.__setattr__(self, "data", tree_or_data)
            object.__setattr__(self, "axes", axes)

    def __setattr__(self, k, v):
        if k in ("data", "axes"):
            object.__setattr__(self, k, v)
        else:
            raise TypeError("assign through .data")

    @staticmethod
    def _build_axes(tree, parts: List[Tensor], offset: int = 0) -> Dict:
        axes: Dict[str, Any] = {}
        start = offset
        for k, v in tree.items():
            if isinstance(v, Tensor):
                n = v.numel()
                axes[k] = ("leaf", offset - start, tuple(v.shape))
                parts.append(_flat_leaf(v))
                offset += n
            else:
                sub_parts: List[Tensor] = []
                sub = ComponentArray._build_axes(v, sub_parts, 0)
                n = sum(p.numel() for p in sub_parts)
                axes[k] = ("tree", offset - start, n, sub)
                parts.extend(sub_parts)
                offset += n
        return axes

    def __len__(self) -> int:
        return int(self.data.numel())

    def keys(self):
        return self.axes.keys()

    def __getattr__(self, k):
        axes = object.__getattribute__(self, "axes")
        key = _lookup(axes, k)
        if key is None:
            raise AttributeError(k)
        ent = axes[key]
        data = object.__getattribute__(self, "data")
        if ent[0] == "leaf":
            _, off, shape = ent
            n = int(np.prod(shape)) if shape else 1
            seg = data[off:off + n]
            return seg.view(shape[1], shape[0]).T if len(shape) == 2 else seg.view(shape)
        _, off, n, sub = ent
        return ComponentArray(data[off:off + n], sub)

    __getitem__ = __getattr__

    def items(self):
        return ((k, getattr(self, k)) for k in self.axes)

    def to_tree(self) -> NT:
        return NT((k, v.to_tree() if isinstance(v, ComponentArray) else v) for k, v in self.items())

    def with_data(self, data: Tensor) -> "ComponentArray":
        return ComponentArray(data, self.axes)

    def requires_grad_(self, flag: bool = True) -> "ComponentArray":
        self.data.requires_grad_(flag)
        return self

    def to(self, device) -> "ComponentArray":
        return ComponentArray(self.data.to(device), self.axes)


def flat_params(ps, n_expected: Optional[int] = None) -> Tensor:
    """Flat float32 parameter segment of a (sub)tree in field order: zero-copy for a ComponentArray, a
    (differentiable) concatenation for a NamedTuple of leaves."""
    if isinstance(ps, ComponentArray):
        flat = ps.data
    else:
        leaves = [_flat_leaf(v) for _, v in _leaves(ps)]
        if len(leaves) == 1:
            flat = leaves[0]
        else:
            flat = torch.cat(leaves) if leaves else torch.zeros(0)
    if n_expected is not None and flat.numel() != n_expected:
        raise ValueError(f"DimensionMismatch: layer expects {n_expected} parameters, got {flat.numel()}")
    return flat
