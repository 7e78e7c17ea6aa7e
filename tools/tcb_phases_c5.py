"""Phase timing of the tensor-core edge-phase backward on the HOISTED C5 VMHConv shape (inner MLP: identity 64x64 + 64 => 64,
dx' = 128): clock64 stamps of CTA 0, worker warp 5."""
import os, sys
# needs the developer build with the stamps compiled in:
#   NGPDE_BUILD_TAG=stamps NGPDE_EXTRA_FLAGS=-DNGPDE_TCB_STAMPS python neuralgraphpde.jl_b200/build.py
os.environ.setdefault("NGPDE_LIB_PATH", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "neuralgraphpde.jl_b200", "libngpde_stamps.so"))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ngpde
from ngpde import _lib, engine, workloads
from ngpde import Chain, Dense, GNNGraph, VMHConv, setup

rng = np.random.default_rng(0)
s1, t1, pos1 = workloads.grid_edges(64, 64, 8, rng)
G = 64
n1 = 64 * 64
offs = (np.arange(G, dtype=np.int64) * n1)[:, None]
s, t = (s1[None, :] + offs).ravel(), (t1[None, :] + offs).ravel()
g = GNNGraph(torch.from_numpy(s), torch.from_numpy(t), num_nodes=n1 * G, num_graphs=G, ndata={"x": torch.from_numpy(np.tile(pos1, (1, G)))}).to("cuda")
h = 64
layer = VMHConv(Chain(Dense(2 * h + 2, h, "tanh"), Dense(h, h)), Chain(Dense(2 * h, h, "tanh"), Dense(h, 2)), initialgraph=g, aggr="mean")
ps, st = setup(rng, layer, "cuda")
x = torch.randn(n1 * G, h, device="cuda").T
r = engine.RhsRunner(layer, x, ps, st)
r.dy.normal_()
for _ in range(3):
    r.step()
buf = torch.zeros(512, dtype=torch.int64, device="cuda:0")
_lib.load().ngpde_debug_buffer(buf.data_ptr())
r.step()
torch.cuda.synchronize()
_lib.load().ngpde_debug_buffer(None)
tt = buf.cpu().view(8, 64)
names = {0: "start", 1: "gather + recompute", 2: "G_L load"}
L = 2
for l in range(L - 1, -1, -1):
    for k, nm in enumerate(["G->TMEM + dgrad issue (from prev end)", "wait prev wgrad", "stage + wgrad issue", "colsum/collect", "wait dgrad"]):
        names[3 + 6 * l + k] = f"L{l} {nm}"
names[28] = "scatter: source-side spill (incl. worker_sync)"
names[29] = "scatter: destination side"
names[30] = "wait layer-0 wgrad batch"
names[27] = "collect dW_0 + end"
order = [0, 1, 2] + [3 + 6 * l + k for l in range(L - 1, -1, -1) for k in range(5)] + [28, 29, 30, 27]
for tile in (2, 5):
    print("tile", tile, "total cycles", int(tt[tile, 27] - tt[tile, 0]))
    prev = int(tt[tile, 0])
    for sl in order[1:]:
        v = int(tt[tile, sl])
        print(f"   {names[sl]:48s} {v - prev:7d}   (t = {v - int(tt[tile, 0])})")
        prev = v
print("tile starts:", [int(tt[i, 0] - tt[0, 0]) for i in range(8)])
