"""Phase timing of the tensor-core edge-phase backward (clock64 stamps of CTA 0, thread 64), C3 workload."""
import os, sys
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
names = {0: "start", 1: "recompute", 2: "G_L"}
for l in (3, 2, 1, 0):
    for k, nm in enumerate(["A:collect", "B:stage0+sync", "D:wait wgrad0", "D:stage1+sync", "F:wait dgrad", "F:next G"]):
        names[3 + 6 * l + k] = f"L{l} {nm}"
names[27] = "dz0 scatter + end"
order = [0, 1, 2] + [3 + 6 * l + k for l in (3, 2, 1, 0) for k in range(6)] + [27]
for tile in (3,):
    print("tile", tile, "total cycles", int(t[tile, 27] - t[tile, 0]))
    prev = int(t[tile, 0])
    for sl in order[1:]:
        v = int(t[tile, sl])
        print(f"   {names[sl]:22s} {v - prev:7d}")
        prev = v

t3 = t[3]
print("recompute layer 1: reach barrier (tid64 idle until tid0's wait returns)", int(t3[34] - t3[0]), "| sync", int(t3[35] - t3[34]), "| ld D + wait", int(t3[36] - t3[35]),
      "| act+split+st issue", int(t3[37] - t3[36]), "| wait::st", int(t3[38] - t3[37]), "| fence+sync", int(t3[39] - t3[38]))
print("F (l=2): ld issue", int(t3[32] - t3[3 + 6 * 2 + 4]), "| wait::ld", int(t3[33] - t3[32]), "| compute+fence", int(t3[3 + 6 * 2 + 5] - t3[33]))
