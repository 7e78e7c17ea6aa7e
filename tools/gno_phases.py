"""Phase timing of the factored GNOConv backward edge kernel: re-times the C4 step with parts of the kernel left out
(NGPDE_OPT_DEBUG_SKIP bitmask: 1 pullback through T, 2 S rebuild, 4 MLP backward, 8 T staging, 16 MLP recompute; 32 is an experiment, not a skip: T_n read from global memory instead of staged in shared
memory -- measured slower, 27.9 vs 21.3 ms for the edge phase at 250k nodes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngpde
from ngpde import _lib, engine, workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 250000
w = workloads.c4_gno("cuda", n_nodes=n)
r = engine.RhsRunner(w.layer, w.x, w.ps, w.st)
r.dy.copy_(torch.randn_like(r.dy))
for mask in (0, 32, 1, 2, 4, 8, 16, 1 | 8, 1 | 2 | 8, 1 | 2 | 4 | 8, 31):
    _lib.set_option(2, mask)
    for _ in range(2):
        r.step()
    _lib.profile_enable(True)
    for _ in range(3):
        r.step()
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    print(f"skip={mask:2d}  " + "  ".join(f"{k}={v[0] / max(v[1], 1):8.3f} ms" for k, v in prof.items()), flush=True)
_lib.set_option(2, 0)
