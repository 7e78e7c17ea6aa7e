"""Build libngpde.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python neuralgraphpde.jl_b200/build.py [--force] [--verbose]

Objects are rebuilt only when their source (or any header) changed; the FFMA engine's kernel templates are instantiated
in three translation units of their own (ngpde_conv_fwd.cu, ngpde_conv_bwd_edge.cu, ngpde_conv_bwd_node.cu), so a full
build takes about 80 s on 8 cores.  The .so is git-ignored but travels to the GPU
box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
# developer variants: NGPDE_BUILD_TAG=x NGPDE_EXTRA_FLAGS="-DFOO=1" builds libngpde_x.so from objects in build_x/
_TAG = os.environ.get("NGPDE_BUILD_TAG", "")
OUT = os.path.join(HERE, f"libngpde_{_TAG}.so" if _TAG else "libngpde.so")
OBJ_DIR = os.path.join(HERE, f"build_{_TAG}" if _TAG else "build")
SOURCES = ["ngpde_conv_bwd_edge.cu", "ngpde_conv_bwd_node.cu", "ngpde_conv_fwd.cu", "ngpde_graph.cu", "ngpde_conv.cu", "ngpde_tc.cu", "ngpde_gcn.cu", "ngpde_halo.cu", "ngpde_gno.cu", "ngpde_gno_tc.cu", "ngpde_train.cu", "ngpde_dist.cu", "ngpde_ode.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", *ARCH, "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
         "-I", os.path.join(ROOT, "include"), "-I", CSRC, *os.environ.get("NGPDE_EXTRA_FLAGS", "").split()]


def _sha(paths) -> str:
    h = hashlib.sha256()
    for p in paths:
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode())
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _headers():
    hs = [os.path.join(CSRC, n) for n in sorted(os.listdir(CSRC)) if n.endswith((".cuh", ".h"))]
    inc = os.path.join(ROOT, "include")
    return hs + [os.path.join(inc, n) for n in sorted(os.listdir(inc))]


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []
    headers = _headers()

    def compile_one(src: str):
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        stamp = obj + ".sha"
        dig = _sha([path] + headers)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            return obj, False
        cmd = [NVCC, *FLAGS, *extra, "-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        with open(stamp, "w") as f:
            f.write(dig)
        return obj, True

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        res = list(ex.map(compile_one, SOURCES))
    objs = [o for o, _ in res]
    if any(changed for _, changed in res) or not os.path.exists(OUT) or force:
        cmd = [NVCC, "-shared", "-o", OUT, *objs, *ARCH, "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
