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
base = int(t3[0])
for l in (3, 2, 1, 0):
    b = 32 + 6 * l
    f = lambda i: int(t3[i]) - base
    print(f"L{l}: layer top {f(3 + 6 * l)} | dgrad issue {f(b)}..{f(b + 1)} | S2 {f(4 + 6 * l)} | wgrad0 issue {f(b + 2)}..{f(b + 3)} | wgrad0 done(seen) {f(5 + 6 * l)} | "
          f"S3 {f(6 + 6 * l)} | wgrad1 issue {f(b + 4)}..{f(b + 5)} | dgrad done(seen) {f(7 + 6 * l)} | next G done {f(8 + 6 * l)} | dgrad done (issuer polls) {f(56 + l)}")

print("last layer: S3", int(t3[6]) - base, "| before bar_d wait", int(t3[60]) - base, "| lane 0 after wait", int(t3[61]) - base, "| after syncwarp", int(t3[7]) - base)
