"""Phase timing of the tensor-core edge-phase backward (clock64 stamps of CTA 0, thread 64), C3 workload."""
import os, sys
# needs the developer build with the stamps compiled in:
#   NGPDE_BUILD_TAG=stamps NGPDE_EXTRA_FLAGS=-DNGPDE_TCB_STAMPS python neuralgraphpde.jl_b200/build.py
os.environ.setdefault("NGPDE_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neuralgraphpde.jl_b200", "libngpde_stamps.so"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ngpde
from ngpde import _lib, engine, workloads
w = workloads.c3_vmh("cuda:0")
r = engine.RhsRunner(w.layer, w.x, w.ps, w.st)
r.dy.normal_()
for _ in range(3):
    r.step()
buf = torch.zeros(512, dtype=torch.int64, device="cuda:0")
_lib.load().ngpde_debug_buffer(buf.data_ptr())
r.step()
torch.cuda.synchronize()
_lib.load().ngpde_debug_buffer(None)
t = buf.cpu().view(8, 64)
names = {0: "start", 1: "gather + recompute", 2: "G_L load"}
for l in (3, 2, 1, 0):
    for k, nm in enumerate(["G->TMEM + dgrad issue (from prev end)", "wait prev wgrad", "stage + wgrad issue", "colsum/collect", "wait dgrad"]):
        names[3 + 6 * l + k] = f"L{l} {nm}"
names[27] = "dz0 scatter + end"
order = [0, 1, 2] + [3 + 6 * l + k for l in (3, 2, 1, 0) for k in range(5)] + [27]
for tile in (2, 5):
    print("tile", tile, "total cycles", int(t[tile, 27] - t[tile, 0]))
    prev = int(t[tile, 0])
    for sl in order[1:]:
        v = int(t[tile, sl])
        print(f"   {names[sl]:48s} {v - prev:7d}   (t = {v - int(t[tile, 0])})")
        prev = v
print("tile starts:", [int(t[i, 0] - t[0, 0]) for i in range(8)])

for tile in (2, 5):
    base = int(t[tile, 0])
    for l in (3, 2, 1, 0):
        f = lambda i: int(t[tile, i]) - base
        print(f"tile {tile} L{l}: layer top {f(3 + 6 * l)} | dgrad issue {f(32 + 4 * l)}..{f(33 + 4 * l)} | prev wgrad seen done {f(4 + 6 * l)} | "
              f"dgrad done (issuer polls, NGPDE_TCB_OPT=4) {f(48 + l)} | wgrad issue {f(34 + 4 * l)}..{f(35 + 4 * l)} | staged+issued {f(5 + 6 * l)} | collected {f(6 + 6 * l)} | dgrad seen done {f(7 + 6 * l)}")
for tile in (5,):
    base = int(t[tile, 0])
    for l in (3, 2, 1, 0):
        f = lambda i: int(t[tile, i]) - base
        print(f"tile {tile} L{l}: staged+fenced {f(5 + 6 * l)} | colsum done {f(52 + l)} | lane 0 poll returned {f(56 + l)} | after __syncwarp {f(6 + 6 * l)} | arrived {f(7 + 6 * l)}")
for tile in (5,):
    base = int(t[tile, 0])
    f = lambda i: int(t[tile, i]) - base
    print(f"tile {tile} recompute: layer 0 MMAs seen done {f(8)} | layer 1 {f(10)} | layer 2 {f(12)} | recompute end {f(1)}")
