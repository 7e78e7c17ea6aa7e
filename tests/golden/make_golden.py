#!/usr/bin/env python
"""Regenerates the golden fixtures in this directory (CPU only; run from the repo root):

    python tests/golden/make_golden.py

Provenance.  The reference (NeuralGraphPDE.jl) is a Julia package and no Julia toolchain exists in this image, so the
reference itself cannot emit vectors.  Two kinds of fixture are therefore kept:

  * `spectral_kat.npz` -- the reference's OWN known-answer test (test/runtests.jl:153-162 and the doctest at
    src/layers.jl:581-631): SpectralConv(100) applied to sin/cos on the 100-point periodic grid must return cos/-sin.
    The expected outputs here are the ANALYTIC values the reference asserts against, not oracle outputs.
  * `<layer>_<case>.npz` -- for the five north-star layers the reference holds no numeric vectors (shape tests only:
    parity unpinned).  These files freeze the ORACLE's float32 results (inputs, parameters, y, dx,
    flat dps) on small seeded cases so that (a) the oracle cannot drift silently and (b) the CUDA path is compared with
    committed numbers, not only with a checker imported at test time.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ngpde  # noqa: E402
from ngpde import workloads  # noqa: E402
import ngpde_oracle as orc  # noqa: E402
from common import oracle_fwd_bwd  # noqa: E402

CASES = {
    "edgeconv_c1_8x8": ("c1", {"side": 8}),
    "vmh_c3_6x6_h16": ("c3", {"side": 6, "hidden": 16}),
    "vmh_c3_5x5_h64": ("c3", {"side": 5, "hidden": 64}),
    "mppde_c2_3x12_h24": ("c2", {"n_per": 12, "n_graphs": 3, "hidden": 24}),
    "gno_c4_60_c4_h8": ("c4", {"n_nodes": 60, "chs": 4, "hidden": 8}),
    "gcn_vmh_c5_2x5x5_h8": ("c5", {"n_graphs": 2, "side": 5, "hidden": 8}),
}


def build_case(name):
    wl, kw = CASES[name]
    w = workloads.WORKLOADS[wl]("cpu", **kw)
    gen = torch.Generator().manual_seed(2024)
    y0, _, _ = oracle_fwd_bwd(w.layer, w.x, w.ps, w.graph)
    dy = torch.randn(tuple(y0.shape), generator=gen)
    y, dx, dps = oracle_fwd_bwd(w.layer, w.x, w.ps, w.graph, dy)
    return w, dy, y, dx, dps


def main():
    for name in CASES:
        w, dy, y, dx, dps = build_case(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), x=w.x.numpy(), dy=dy.numpy(), y=y.numpy(), dx=dx.numpy(),
                            dps=dps.numpy(), ps=ngpde.ComponentArray(w.ps).data.numpy(), s=w.graph.s.numpy(),
                            t=w.graph.t.numpy())
        print(name, "y", tuple(y.shape), "params", dps.numel())
    n = 100
    xs = np.linspace(0.0, 2.0 * np.pi, n + 1)[1:]  # LinRange(0, 2pi, 101)[2:end], test/runtests.jl:157
    np.savez_compressed(os.path.join(HERE, "spectral_kat.npz"), x=xs, sin=np.sin(xs), cos=np.cos(xs))
    print("spectral_kat", n)


if __name__ == "__main__":
    main()
