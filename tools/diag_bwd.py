"""Per-block error of the backward on one config: tensor-core path and FFMA path, each against the float64 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import ngpde
from ngpde import Chain, Dense, GNNGraph, VMHConv, setup
from common import jl_rand, oracle_fwd_bwd, product_fwd_bwd, relerr

DEV = "cuda:0"
import json
aggr = sys.argv[1] if len(sys.argv) > 1 else "+"
long_row = int(sys.argv[2]) if len(sys.argv) > 2 else 300
dims = json.loads(sys.argv[3]) if len(sys.argv) > 3 else [6, 16, 48, 5]
gdims = json.loads(sys.argv[4]) if len(sys.argv) > 4 else [7, 40, 3]
rng = np.random.default_rng(77)
n = 900
s, t = rng.integers(0, n, 7000), rng.integers(0, n, 7000)
t[:long_row] = 11
t[300:310] = n - 1
dx = dims[0] // 2 - 1
g = GNNGraph(torch.from_numpy(s), torch.from_numpy(t), num_nodes=n, ndata={"x": jl_rand(rng, 2, n)}).to(DEV)
acts = ["tanh", "sigmoid", "elu", "softplus"]
mk = lambda dd: Chain(*[Dense(dd[i], dd[i + 1], acts[i % 4] if i < len(dd) - 2 else "identity") for i in range(len(dd) - 1)])
layer = VMHConv(mk(dims), mk(gdims), initialgraph=g, aggr=aggr)
ps, st = setup(rng, layer, DEV)
x = jl_rand(rng, dx, n, DEV)
r2 = np.random.default_rng(123)
y0, _, _ = product_fwd_bwd(layer, x, ps, st)
dy = torch.from_numpy(r2.standard_normal(tuple(y0.shape)).astype(np.float32)).to(DEV)
y64, dx64, dp64 = oracle_fwd_bwd(layer, x, ps, g, dy, torch.float64)
y32, dx32, dp32 = oracle_fwd_bwd(layer, x, ps, g, dy, torch.float32)
blocks, off = [], 0
for nm, dd in (("phi", dims), ("gam", gdims)):
    for i in range(len(dd) - 1):
        blocks.append((f"{nm}.W{i}", off, off + dd[i] * dd[i + 1])); off += dd[i] * dd[i + 1]
        blocks.append((f"{nm}.b{i}", off, off + dd[i + 1])); off += dd[i + 1]
for name, tc in (("TC", 1), ("FFMA", 0)):
    ngpde._lib.set_option(ngpde._lib.OPT_TENSOR_CORES, tc)
    y, dxx, dp = product_fwd_bwd(layer, x, ps, st, dy)
    print(name, "y", relerr(y, y64), "dx", relerr(dxx, dx64), "dp", relerr(dp, dp64), "| oracle32: dx", relerr(dx32, dx64), "dp", relerr(dp32, dp64))
    for nm, a, b in blocks:
        print(f"   {nm:8s} rel {relerr(dp[a:b], dp64[a:b]):.2e}  max|ref| {dp64[a:b].abs().max().item():.3e}")
    for r in range(dxx.shape[0]):
        print(f"   dx[{r}] rel {relerr(dxx[r], dx64[r]):.2e} (oracle32 {relerr(dx32[r], dx64[r]):.2e})  max|ref| {dx64[r].abs().max().item():.3e}")
    e = (dxx.cpu().double() - dx64).abs()
    i = int(e.max(dim=0).values.argmax())
    print("   dx worst node", i, "err", e[:, i].tolist(), "ref", dx64[:, i].tolist(), "indeg", int((t == i).sum()), "outdeg", int((s == i).sum()))
