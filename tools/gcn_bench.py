"""GCN aggregate kernel variants (NGPDE_GCN_V) on the C5 graph (512 x 64x64 grid-8 = 2M nodes), GCNConv(64 => 64) forward.
Run under  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:gcn_aggregate  for the
kernel's own duration / DRAM bytes; the printed CUDA-event time is the whole layer forward."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import ngpde
from ngpde import workloads
from ngpde.layers import GCNConv
from ngpde.lux import setup

graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 512
w = workloads.c5_gcn_vmh("cuda", n_graphs=graphs)
g = w.graph
rng = np.random.default_rng(1)
layer = GCNConv((64, 64), "tanh", initialgraph=g)
ps, st = setup(rng, layer, "cuda")
x = torch.randn(g.num_nodes, 64, device="cuda").T   # Julia shape (64, N) over row-major [N][64] storage: no layout copy
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ref = None
variants = [int(a) for a in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 1, 2]
for v in variants:
    os.environ["NGPDE_GCN_V"] = str(v)
    y, _ = layer(x, ps, st)
    torch.cuda.synchronize()
    if ref is None:
        ref = y.clone()
    same = torch.equal(ref, y)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        layer(x, ps, st)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"NGPDE_GCN_V={v}: layer forward {min(ts) * 1e3:.1f} us (min of 5, L2 flushed), bit-identical to variant 0: {same}")
