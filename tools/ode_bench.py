"""C1 Neural-ODE trajectory (Tsit5, 20 steps, 120 RHS) forward + adjoint: eager autograd path, CUDA-graph step path, persistent kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import ngpde
from ngpde import ode, workloads

w = workloads.c1_edgeconv("cuda")
dt, nsteps = 0.05, 20
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
rk = ode.GraphedRK(w.layer, w.x, w.ps, w.st, dt, "tsit5")
prk = ode.PersistentRK(w.layer, w.x, w.ps, w.st, dt, "tsit5")
g = torch.ones_like(rk.u)
def graphed():
    rk.solve(w.x, nsteps); rk.adjoint(g)
def persistent():
    prk.solve(w.x, nsteps); prk.adjoint(g)
def persistent_fwd():
    prk.solve(w.x, nsteps)
rhs = nsteps * 6
tg, tp, tpf = timeit(graphed), timeit(persistent), timeit(persistent_fwd)
print(f"C1 Tsit5 {nsteps} steps ({rhs} RHS): CUDA-graph step path fwd+adjoint {tg:.3f} ms ({1e3 * tg / rhs:.2f} us/RHS, 3 RHS-equivalents each); "
      f"persistent kernels fwd+adjoint {tp:.3f} ms ({1e3 * tp / rhs:.2f} us/RHS); forward only {tpf:.3f} ms ({1e3 * tpf / rhs:.2f} us/RHS)")

# phase breakdown (cycles of CTA 0 thread 0, summed over the trajectory)
from ngpde import _lib
buf = torch.zeros(512, dtype=torch.int64, device="cuda:0")
_lib.load().ngpde_debug_buffer(buf.data_ptr())
prk.solve(w.x, nsteps); prk.adjoint(g)
torch.cuda.synchronize()
_lib.load().ngpde_debug_buffer(None)
t = buf.cpu().tolist()
print("forward, cycles per RHS:", dict(zip(["edges", "nodes+dsmem", "cluster.sync"], [round(v / rhs) for v in t[0:3]])))
print("forward, warp 0 (sum over its tiles) cycles per RHS:", dict(zip(["gather", "L0", "L1", "L2", "L3"], [round(v / rhs) for v in t[3:8]])))
print("adjoint, cycles per RHS:", dict(zip(["-", "tiles: recompute + backprop + dW", "-", "cluster.sync", "ubar + next kbar"], [round(v / rhs) for v in t[8:13]])))
