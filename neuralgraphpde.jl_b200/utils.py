"""Graph-in-state utilities (/root/reference/src/utils.jl:1-31)."""
from __future__ import annotations

from typing import Optional

from .graph import GNNGraph, copy
from .lux import NT


def drop(nt: dict, key: str) -> NT:
    """drop(nt, key) = Base.structdiff(nt, NamedTuple{(key,)})  (utils.jl:1)."""
    return NT((k, v) for k, v in nt.items() if k != key)


def updategraph(st: NT, g: Optional[GNNGraph] = None, **kwargs) -> NT:
    """Recursively replace every `graph` leaf of the state tree by `g` (arrays shared), or -- when `g` is None --
    by `copy(old; kwargs...)` to swap only ndata/edata/gdata (utils.jl:24-31).

    This is the only point at which topology changes, so it is also where the cached CSR layout is invalidated:
    a new GNNGraph brings its own (lazily built) libngpde handle, a data-only update keeps sharing the old one.
    The reference asserts `updategraph(st, g).graph === g` (test/runtests.jl:179,184): the graph object itself is
    installed, matching GNNGraph's `===` on immutable structs with identical fields.
    """
    if len(st) == 0:
        return st
    out = NT()
    for k, v in st.items():
        if isinstance(v, GNNGraph):
            new = g if g is not None else copy(v, **kwargs)
            dict.__setitem__(out, k, new)
        elif isinstance(v, dict):
            dict.__setitem__(out, k, updategraph(NT(v), g, **kwargs))
        else:
            dict.__setitem__(out, k, v)
    return out
