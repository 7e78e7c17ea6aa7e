#!/usr/bin/env python
"""Benchmark of the message-passing hot path (BASELINE.json metric: edge-messages/s and ODE-RHS-evals/s, fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one right-hand-side evaluation, forward plus its vector-Jacobian product w.r.t. (x, ps), of the workload's
layer over its synthetic graph -- by default C3: VMHConv on the 256x256 grid-8 graph (65,536 nodes, 521,220 edges,
hidden 64), the configuration BASELINE.json's target is quoted on.  With N > 1 ranks every rank owns one such graph
(batched-ensemble sharding, SURVEY.md 8e: no data-path collective; the flat parameter gradient is all-reduced over NCCL
every step, as a data-parallel training step does), so scaling is weak.

One JSON line on stdout (rank 0).  `value` = edge messages (fwd+bwd) per second over all ranks with inputs resident in
HBM; `e2e` = the same through the public layer API with pinned HOST buffers copied in and results copied out every
step; `roofline` = the dominant kernel (edge-phase backward) timed with CUDA events inside the library;
`cpu_baseline` = the oracle's unfused restatement of the reference algorithm on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np
import torch

METRIC = "edge_messages_per_sec_fwd_bwd"
UNIT = "edge-messages/s"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def lib_sha256() -> str:
    import hashlib
    path = os.path.join(ROOT, "neuralgraphpde.jl_b200", "libngpde.so")
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for chunk in iter(lambda: f.read(1 << 20), b""):
            h.update(chunk)
    return h.hexdigest()


def committed_traffic(workload: str, kernel: str):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/traffic.json), valid only for
    the build it was captured on: the file records the sha256 of libngpde.so and a stale capture is reported as null."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return None, "no capture committed"
    with open(tpath) as f:
        t = json.load(f)
    sha = lib_sha256()
    if t.get("lib_sha256") != sha:
        return None, f"capture is of another build (captured {str(t.get('lib_sha256'))[:12]}, benched {sha[:12]})"
    return t.get(workload, {}).get(kernel), f"profiles/traffic.json (ncu --set full, build {sha[:12]})"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.05):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.sm, self.reasons, self.power = [], set(), []
        self.sm_max = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    _NAMES = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
              0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
              0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                try:
                    r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:  # noqa: BLE001
                    r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self._NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "power_w_max": max(self.power) if self.power else None, "samples": len(self.sm)}


def physical_gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:  # noqa: BLE001
            return local_rank
    return local_rank


def cpu_reference_step_fn(workload_name: str, sample_kw: dict):
    """The oracle's (unfused, torch-CPU float32) forward + backward on a bounded sample of the workload."""
    import ngpde
    from ngpde import workloads
    from common import oracle_forward, to_ograph, tree_requires_grad, tree_to_cpu
    w = workloads.WORKLOADS[workload_name]("cpu", **sample_kw)
    og = to_ograph(w.graph)
    pc = tree_requires_grad(tree_to_cpu(w.ps))
    x = w.x.detach().clone().contiguous().requires_grad_(True)
    gen = torch.Generator().manual_seed(0)
    y0 = oracle_forward(w.layer, x, pc, og)
    dy = torch.randn(y0.shape, generator=gen)

    def step():
        x.grad = None
        y = oracle_forward(w.layer, x, pc, og)
        y.backward(dy)
        return y

    return step, w


# bounded CPU samples of each workload ({} = the workload itself, at full size: C3 is 0.2-0.3 s per fwd+bwd on 16 cores)
CPU_SAMPLE = {"c1": {}, "c2": {}, "c3": {}, "c4": {"n_nodes": 15625}, "c5": {"n_graphs": 8}}


def run_reference(args, rank: int):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    sample_kw = CPU_SAMPLE[args.workload]
    step, w = cpu_reference_step_fn(args.workload, sample_kw)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = w.n_edges * args.steps / dt
    sample = (f"{w.name}: {w.n_nodes} nodes / {w.n_edges} edges per step "
              f"({'the full workload' if not sample_kw else 'reduced: ' + str(sample_kw)})")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(args.workload), "sample": sample,
                   "note": "CPU restatement (torch) of the reference's unfused gather->Dense->scatter algorithm; the Julia "
                           "package itself cannot run in this image"},
        "rhs_evals_per_sec": args.steps / dt,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_label(name: str) -> str:
    return {
        "c1": "C1 ExplicitEdgeConv Neural-ODE RHS, 32x32 grid-4 graph (1,024 nodes / 3,968 edges), hidden 16",
        "c2": "C2 MPPDEConv, 64 x 256-node path graphs batched (16,384 nodes / 32,640 edges), hidden 128",
        "c3": "C3 VMHConv RHS, 256x256 grid-8 graph (65,536 nodes / 521,220 edges), phi 6-64-64-64-64, gamma 66-64-64-64-2",
        "c4": "C4 GNOConv 64=>64, random-geometric graph 1M nodes / ~16M edges, phi 6-64-64-4096",
        "c5": "C5 GCNConv+GCNConv+VMHConv RHS on 512 x (64x64 grid-8) graphs",
    }[name]


class Timer:
    """K steps bracketed by barrier + synchronize on both sides, every step timed with CUDA events on the launching
    stream, L2 flushed (untimed) between steps; the result is the MAX over ranks of the summed step times."""

    def __init__(self, dev, world, flush):
        self.dev, self.world, self.flush = dev, world, flush

    def sync_all(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(self.dev)

    def __call__(self, fn, k):
        evs = []
        self.sync_all()
        for _ in range(k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
            if self.flush is not None:
                self.flush.zero_()
        self.sync_all()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms


def roofline_block(workload, w, prof, K, paths, step_ms, peaks, edges_launch, nodes_launch):
    """`roofline` object for the dominant fused kernel.  `edges_launch` / `nodes_launch` are the units ONE launch on THIS
    rank processes (a node-partitioned rank holds 1/world of the graph: its flops, not the global graph's)."""
    dom = max(prof, key=lambda k: prof[k][0])
    dom_ms = prof[dom][0] / max(prof[dom][1], 1)
    phase_e = dom.endswith("edge")
    fe, fn = w.notes["flops_edge_fwd"] / max(w.n_edges, 1), w.notes["flops_node_fwd"] / max(w.n_nodes, 1)
    f_phase = fe * edges_launch if phase_e else fn * nodes_launch
    # algorithmic flops of that launch: forward F; backward 2F (dgrad + wgrad); recompute not counted
    alg_flops = f_phase * (1.0 if dom.startswith("fwd") else 2.0)
    ach_tflops = alg_flops / (dom_ms * 1e-3) / 1e12
    step_kernel_ms = sum(v[0] for v in prof.values()) / K
    on_tc = paths.get(dom, 0) in (1, 3)
    factored = paths.get(dom, 0) == 2
    kname = {"fwd_edge": "mp_fwd{}_kernel<edge>", "fwd_node": "mp_fwd{}_kernel<node>", "bwd_node": "mp_bwd{}_kernel<node>",
             "bwd_edge": "mp_bwd{}_kernel<edge>"}[dom].format("_tc" if on_tc else "")
    if paths.get(dom, 0) == 3:
        kname = ("%s phase, one gno_gemm_tc_kernel (tcgen05 3xTF32) per Dense layer + elementwise kernels "
                 "(csrc/ngpde_layered.cuh)" % dom)
    if factored:
        kname = {"fwd_edge": "gno factored edge phase: phi hidden layers (gno_gemm_tc GEMMs, or mp_fwd_kernel<edge> with "
                             "NGPDE_OPT_GNO_LAYERED=0) + gno_node_kernel (S builder) + gno_gemm (mbar = S B)",
                 "bwd_edge": "gno factored edge phase: gno_gemm (T = DM B') + gno_node_kernel (dz, dh per destination) + phi hidden "
                             "layers' backward GEMMs + gno_gemm (dB = S' DM)"}[dom]
    # the pipe that can hold the 1e-5 tolerance: 3xTF32 on tcgen05 = a third of the TF32 rate = a sixth of the dense BF16
    # peak; the FP32-FFMA engine: 148 SMs x 128 lanes x 2 flop x 1.965 GHz = 74.4 TFLOP/s
    pipe_peak = peaks["bf16_tflops"] / 6.0 if on_tc else 74.4
    traffic, traffic_src = committed_traffic(workload, kname)
    scale = edges_launch / max(w.n_edges, 1)
    ex = w.notes.get("flops_executed_" + dom)
    out = {
        "kernel": kname,
        "bound": "tensor", "achieved": ach_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
        "frac": ach_tflops / peaks["bf16_tflops"], "traffic": traffic, "traffic_source": traffic_src,
        "peak_source": peaks["source"] + " bf16 cuBLAS burst (MEASURED_PEAKS.json)",
        "pipe": ("tcgen05 3xTF32 (fp32-accurate split: 3 TF32 MMAs per product; peak = bf16 peak / 6)" if on_tc else
                 ("mixed: tcgen05 3xTF32 GEMMs (phi hidden layers, T, mbar, dB) + fp32 FFMA per-destination kernel; fractions "
                  "quoted against the fp32 FFMA peak" if factored else
                  "fp32 FFMA (fp32-accurate; tolerance 1e-5 excludes plain TF32/BF16)")),
        "pipe_peak_tflops": pipe_peak, "pipe_frac": ach_tflops / pipe_peak,
        "kernel_paths": {k: {1: "tcgen05", 0: "ffma", 2: "factored (ffma + fp32 gemm)", 3: "tcgen05 GEMM per Dense layer", -1: "none"}[v] for k, v in paths.items()},
        "avg_launch_ms": dom_ms, "algorithmic_flops_per_launch": alg_flops,
        "units_per_launch": {"edges": int(edges_launch), "nodes": int(nodes_launch)},
        "share_of_step_kernel_time": (prof[dom][0] / K) / step_kernel_ms if step_kernel_ms else None,
        "kernels_ms_per_step": {k: v[0] / K for k, v in prof.items()},
        "hbm": {"algorithmic_bytes_per_step": w.bytes_fwdbwd * scale,
                "achieved_gbs": w.bytes_fwdbwd * scale / (step_ms * 1e-3) / 1e9,
                "frac_of_measured": w.bytes_fwdbwd * scale / (step_ms * 1e-3) / 1e9 / peaks["hbm_gbs"]},
    }
    if factored and ex is not None:
        # the factored evaluation does fewer flops than SURVEY 8d's count: lead with what is executed
        ex = ex * scale
        out.update({
            "executed_flops_per_launch": ex, "executed_tflops": ex / (dom_ms * 1e-3) / 1e12,
            "executed_pipe_frac": ex / (dom_ms * 1e-3) / 1e12 / pipe_peak,
            "note": "factored GNOConv (csrc/ngpde_gno.cuh): the contraction with phi's affine last layer is done once per "
                    "destination node on per-node outer-product sums instead of once per edge, so the flops EXECUTED are "
                    "fewer than SURVEY 8d's algorithmic count. `executed_*` is the honest pipe utilisation; `achieved` / "
                    "`pipe_frac` use the algorithmic count as the contract asks (and can exceed the pipe: work not done, "
                    "not work done faster)"})
    return out


def ode_trajectory_c1(w, dt=0.05, nsteps=20, reps=20):
    """Tsit5, fixed step, 20 steps = 120 right-hand sides (docs/src/tutorials/graph_node.md:53-66), forward + discrete adjoint:
    the CUDA-graph-captured step built from the layer kernels (ode.GraphedRK) against the two persistent cluster kernels
    (ode.PersistentRK: ngpde_edgeconv_ode_forward / _adjoint, ONE launch each)."""
    from ngpde import ode
    rk = ode.GraphedRK(w.layer, w.x, w.ps, w.st, dt, "tsit5")
    prk = ode.PersistentRK(w.layer, w.x, w.ps, w.st, dt, "tsit5")
    g = torch.ones_like(rk.u)
    rhs = nsteps * 6

    def timeit(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    t_graph = timeit(lambda: (rk.solve(w.x, nsteps), rk.adjoint(g)))
    t_pers = timeit(lambda: (prk.solve(w.x, nsteps), prk.adjoint(g)))
    t_pers_fwd = timeit(lambda: prk.solve(w.x, nsteps))
    uT = prk.solve(w.x, nsteps).clone()
    uT2 = rk.solve(w.x, nsteps)
    return {"method": "tsit5, dt 0.05, 20 steps (120 right-hand sides)", "rhs_per_trajectory": rhs,
            "cuda_graph_step_path_us_per_rhs_fwd_adjoint": 1e3 * t_graph / rhs,
            "persistent_kernels_us_per_rhs_fwd_adjoint": 1e3 * t_pers / rhs,
            "persistent_kernels_us_per_rhs_fwd": 1e3 * t_pers_fwd / rhs,
            "persistent_rhs_evals_per_sec_fwd_adjoint": rhs / (t_pers * 1e-3),
            "launches_per_trajectory_persistent": 2,
            "max_abs_diff_between_paths": float((uT - uT2).abs().max())}


def strong_c4(args, rank, world, dev, peaks):
    """The north-star scaling configuration: ONE GNOConv graph of 1M nodes / ~16M radius edges, node-partitioned over
    the `world` ranks with a halo exchange per RHS, fwd + VJP per step, strong scaling; with a parity
    record computed on the spot: this rank's owned rows against the same call on the unpartitioned graph."""
    import torch.distributed as dist
    from ngpde import _lib, distributed as D, engine, ops, workloads
    n_nodes = args.c4_nodes
    K = max(3, min(args.steps, args.c4_steps))
    t_build = time.perf_counter()
    w = workloads.c4_gno(dev, n_nodes=n_nodes)
    gen = torch.Generator().manual_seed(4321)
    dy_full = torch.randn((w.layer.out_chs, w.n_nodes), generator=gen).to(dev)
    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    timer = Timer(dev, world, flush)
    out = {"workload": workload_label("c4"), "nodes_total": w.n_nodes, "edges_total": w.n_edges, "steps": K,
           "scaling": "strong", "unit": UNIT}
    if world == 1:
        runner = engine.RhsRunner(w.layer, w.x, w.ps, w.st)
        runner.dy.copy_(dy_full.T)
        step = runner.step
        edges_launch, nodes_launch = w.n_edges, w.n_nodes
        out["parallelism"] = "single GPU (the 1-GPU point of the strong-scaling series)"
    else:
        pl = D.PartitionedLayer(w.layer, w.graph, rank, world, dev, mode=args.halo)
        pr = engine.PartitionedRhsRunner(pl, pl.owned(w.x), w.ps, w.st)
        pr.dy_owned.copy_(pl.owned(dy_full).T)
        runner, step = pr.runner, pr.step
        p = pl.part
        edges_launch, nodes_launch = int(p.edge_ids.size), int(p.n_local)
        halo = torch.tensor([p.n_halo, p.n_owned, p.edge_ids.size], dtype=torch.float64, device=dev)
        hmax = halo.clone()
        dist.all_reduce(hmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(halo, op=dist.ReduceOp.SUM)
        out.update({"parallelism": f"node-partitioned over {world} GPUs (contiguous ranges balanced by in-edges), halo via {args.halo}",
                    "halo_rows_max": int(hmax[0].item()), "halo_rows_total": int(halo[0].item()),
                    "owned_rows_max": int(hmax[1].item()), "edges_max_over_mean": float(hmax[2].item() / (halo[2].item() / world)),
                    "step": "halo exchange + layer forward + VJP + reverse halo + dW all-reduce"})
    out["build_s"] = time.perf_counter() - t_build
    for _ in range(3):
        step()
        if flush is not None:
            flush.zero_()
    l0 = ops.LAUNCHES["count"]
    ms = timer(step, K)
    out.update({"ms_per_step": ms / K, "value": w.n_edges * K / (ms * 1e-3), "rhs_evals_per_sec": K / (ms * 1e-3),
                "gpu_launches": ops.LAUNCHES["count"] - l0})
    _lib.profile_enable(True)
    timer(step, K)
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    out["roofline"] = roofline_block("c4", w, prof, K, _lib.kernel_paths(runner.handle, runner.desc), ms / K, peaks,
                                     edges_launch, nodes_launch)
    # ---- parity, at this world size, on the spot ----
    if world > 1:
        step()
        y_p, dx_p = pr.y_owned.clone(), pr.dx_owned.clone()
        dp_p = runner.dparams.clone()
        del pr, runner, step   # (`step` is the partitioned runner's bound method: it keeps its workspaces alive)
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        # the same call on the whole graph, on this rank's GPU; without the forward->backward state buffer (31 GB at 1M nodes):
        # the backward recomputes instead, bit-identical results (tests/test_gpu_parity.py)
        full = engine.RhsRunner(w.layer, w.x, w.ps, w.st, use_state=False)
        full.dy.copy_(dy_full.T)
        full.step()
        own = pl.owned_global.to(dev)
        den = lambda t: max(float(t.abs().max().item()), 1e-30)
        mism = float((y_p != full.y[own]).sum().item())
        e_dx = float((dx_p - full.dx[own]).abs().max().item()) / den(full.dx)
        e_dp = float((dp_p - full.dparams).abs().max().item()) / den(full.dparams)
        stats = torch.tensor([mism, e_dx, e_dp], dtype=torch.float64, device=dev)
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        out["dist_parity"] = {"ranks": world, "forward_mismatched_values": int(stats[0].item()),
                              "dx_rel_err": float(stats[1].item()), "dparams_rel_err": float(stats[2].item()),
                              "against": "the unpartitioned layer call (forward + VJP) on the same 1M-node graph, run on every "
                                         "rank's own GPU; forward must be bit-identical, gradients within 1e-5 (max over ranks)",
                              "ok": bool(stats[0].item() == 0 and stats[1].item() <= 1e-5 and stats[2].item() <= 1e-5)}
        del full
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--partition", action="store_true",
                    help="N>1: split ONE graph over the ranks by nodes with a per-step halo exchange (strong scaling) "
                         "instead of one graph per rank (weak scaling)")
    ap.add_argument("--halo", default="nccl", choices=["nccl", "put", "native"],
                    help="halo transfer: NCCL all-to-all (torch), peer stores, or the C ABI's own NCCL communicator")
    ap.add_argument("--nodes", type=int, default=0, help="override the node count of c4 (default 1,000,000)")
    ap.add_argument("--cuda-graph", dest="cuda_graph", action="store_true", default=True,
                    help="replay the step's forward+backward launches from a captured CUDA graph (default)")
    ap.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false", help="launch every kernel directly")
    ap.add_argument("--graphs", type=int, default=0, help="c5: total number of 64x64 graphs in the ensemble (default 512)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="keep L2 warm between steps (reported under config)")
    ap.add_argument("--allreduce", default="nccl", choices=["peer", "nccl"],
                    help="N>1, one graph per GPU: the gradient sum over ranks -- 'peer': one-shot kernel over peer-mapped "
                         "buffers (NVLink/NVSwitch, ngpde_peer_allreduce_sum), 'nccl': torch.distributed all_reduce")
    ap.add_argument("--no-strong-c4", action="store_true",
                    help="skip the node-partitioned C4 (1M-node GNOConv) strong-scaling block that the default C3 line carries")
    ap.add_argument("--c4-nodes", type=int, default=1_000_000)
    ap.add_argument("--c4-steps", type=int, default=10)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    import ngpde
    from ngpde import _lib, engine, ops, workloads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the message-passing path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()

    wkw = {"n_nodes": args.nodes} if (args.workload == "c4" and args.nodes) else {}
    chain = args.workload == "c5"
    if chain:
        # the ensemble config: 512 graphs in total, whole graphs per rank (no data-path collective, dW all-reduce)
        wkw = {"n_graphs": max(1, (args.graphs or 512) // world)}
    w = workloads.WORKLOADS[args.workload](dev, **wkw)
    K, W = args.steps, args.warmup
    gen = torch.Generator().manual_seed(1234)
    partitioned = args.partition and world > 1
    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    timed = Timer(dev, world, flush)
    allreduce_in_graph = False
    if partitioned:
        from ngpde import distributed as D
        pl = D.PartitionedLayer(w.layer, w.graph, rank, world, dev, mode=args.halo)
        prunner = engine.PartitionedRhsRunner(pl, pl.owned(w.x), w.ps, w.st)
        runner = prunner.runner
        prunner.dy_owned.copy_(torch.randn(tuple(prunner.dy_owned.shape), generator=gen).to(dev))
        one_step = prunner.step
    elif chain:
        runner = engine.ChainRhsRunner(w.layer, w.x, w.ps, w.st)
        runner.dy.copy_(torch.randn(tuple(runner.dy.shape), generator=gen).to(dev))

        def one_step():
            runner.step()
            if world > 1:
                dist.all_reduce(runner.dparams)
    else:
        # data-parallel: the flat [dphi | dnode] gradient is summed over ranks once per step.  Default: the backward writes it
        # into a peer-mapped symmetric buffer and ONE kernel per rank adds all copies over NVLink in rank order (deterministic,
        # identical on every rank, no NCCL launch latency); --allreduce nccl: one coalesced NCCL all-reduce.  Either way it is
        # captured into the same CUDA graph as the kernels when the capture accepts it.
        par = None
        if world > 1 and args.allreduce == "peer":
            try:
                from ngpde import distributed as D
                par = D.PeerAllReduce(engine.RhsRunner.dparams_len(w.layer, w.x, w.ps, w.st), dev)
            except Exception as exc:  # noqa: BLE001
                sys.stderr.write(f"bench.py: peer-memory all-reduce unavailable ({exc!r}); using NCCL\n")
                par = None
        # every rank must take the same path
        if world > 1 and args.allreduce == "peer":
            ok = torch.tensor([1 if par is not None else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                par = None
        runner = engine.RhsRunner(w.layer, w.x, w.ps, w.st, dparams=None if par is None else par.buffer)
        runner.dy.copy_(torch.randn(tuple(runner.dy.shape), generator=gen).to(dev))

        def grad_allreduce():
            if par is not None:
                par.reduce()
            else:
                dist.all_reduce(runner.dparams)

        if args.cuda_graph:
            try:
                # the peer-memory path is issued AFTER the replay: its cross-GPU barrier (symmetric-memory signal pads) must
                # not be frozen into a captured graph (a replayed barrier dead-locked in testing)
                in_graph = world > 1 and par is None
                runner.capture(extra=grad_allreduce if in_graph else None)
                allreduce_in_graph = in_graph
            except Exception as exc:  # noqa: BLE001
                torch.cuda.synchronize(dev)
                try:
                    runner.capture()
                except Exception as exc2:  # keep measuring with direct launches rather than lose the line
                    sys.stderr.write(f"bench.py: CUDA graph capture failed ({exc2}); launching kernels directly\n")
                    runner.graph = None
                    args.cuda_graph = False
                    torch.cuda.synchronize(dev)
                else:
                    sys.stderr.write(f"bench.py: all-reduce not capturable ({exc}); it is issued after the graph replay\n")

        def one_step():
            runner.step()
            if world > 1 and not allreduce_in_graph:
                grad_allreduce()

    for _ in range(W):
        one_step()
        if flush is not None:
            flush.zero_()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    l0 = ops.LAUNCHES["count"]
    total_ms = timed(one_step, K)
    launches = ops.LAUNCHES["count"] - l0
    clocks = sampler.stop()

    edges_total = w.n_edges * K * (1 if partitioned else world)
    value = edges_total / (total_ms * 1e-3)

    # ---- dominant kernel, timed by the library's own CUDA events on the launching stream ----
    _lib.profile_enable(True)
    if args.cuda_graph and not partitioned and not chain:
        # a graph replay does not pass through the library's event hooks: time the same kernels launched directly
        def prof_step():
            runner.forward()
            runner.backward()
        timed(prof_step, K)
    else:
        timed(one_step, K)
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    paths = _lib.kernel_paths(runner.handle, runner.desc)
    if partitioned:
        edges_launch, nodes_launch = int(pl.part.edge_ids.size), int(pl.part.n_local)
    else:
        edges_launch, nodes_launch = w.n_edges, w.n_nodes
    roofline = roofline_block(args.workload, w, prof, K, paths, total_ms / K, peaks, edges_launch, nodes_launch)

    if partitioned:
        p = pl.part
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_label(args.workload), "nodes_total": w.n_nodes, "edges_total": w.n_edges,
                       "nodes_owned_rank0": p.n_owned, "halo_rows_rank0": p.n_halo, "edges_rank0": int(p.edge_ids.size),
                       "step": "one RHS evaluation: halo exchange + layer forward + VJP + reverse halo + dW all-reduce",
                       "parallelism": f"node-partitioned over {world} GPUs, halo via {args.halo}",
                       "l2": "warm (--no-flush)" if flush is None else "flushed between steps (256 MiB memset, untimed)"},
            "rhs_evals_per_sec": K / (total_ms * 1e-3), "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        }
        if rank == 0:
            print(json.dumps(line), flush=True)
        dist.destroy_process_group()
        return

    # ---- end to end through the public layer API with host buffers ----
    ca = ngpde.ComponentArray(w.ps)
    ca.data.requires_grad_(True)
    x_host = w.x.T.contiguous().cpu().pin_memory()          # [N, d] == Julia (d, N)
    dy_host = runner.dy.cpu().pin_memory()
    x_dev = torch.empty_like(x_host, device=dev)
    dy_dev = torch.empty_like(dy_host, device=dev)
    y_host = torch.empty(tuple(runner.y.shape), dtype=torch.float32).pin_memory()
    dx_host = torch.empty_like(x_host).pin_memory()
    dp_host = torch.empty(ca.data.numel(), dtype=torch.float32).pin_memory()
    h2d = x_host.numel() * 4 + dy_host.numel() * 4
    d2h = y_host.numel() * 4 + dx_host.numel() * 4 + dp_host.numel() * 4
    # the resident-step runner's workspaces and forward->backward state are not needed any more; the public-API leg allocates
    # its own (C4 at 1M nodes: 63 GB + 33 GB each -- both sets do not fit 180 GB at once)
    graph_kernels = getattr(runner, "graph_kernels", None)
    for name in ("graph", "ws", "ws_fwd", "state"):
        if hasattr(runner, name):
            setattr(runner, name, None)
    torch.cuda.empty_cache()

    def e2e_step():
        x_dev.copy_(x_host, non_blocking=True)
        dy_dev.copy_(dy_host, non_blocking=True)
        xin = x_dev.T.requires_grad_(True)
        ca.data.grad = None
        y, _ = w.layer(xin, ca, w.st)
        y.backward(dy_dev.T)
        if world > 1:
            dist.all_reduce(ca.data.grad)
        y_host.copy_(y.detach().T, non_blocking=True)
        dx_host.copy_(xin.grad.T, non_blocking=True)
        dp_host.copy_(ca.data.grad, non_blocking=True)

    for _ in range(W):
        e2e_step()
    e2e_ms = timed(e2e_step, K)
    e2e_value = edges_total / (e2e_ms * 1e-3)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong" if chain else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(args.workload), "nodes_per_gpu": w.n_nodes, "edges_per_gpu": w.n_edges,
                   **({"graphs_per_gpu": wkw["n_graphs"], "graphs_total": wkw["n_graphs"] * world,
                       "chain": "GCNConv(2=>64,tanh) -> GCNConv(64=>64,tanh) -> VMHConv; the roofline block covers the "
                                "VMHConv kernels only (the GCN aggregate is HBM-bound: DESIGN.md section 4)"} if chain else {}),
                   "step": "one RHS evaluation: layer forward + VJP w.r.t. (x, ps)", "aggr": "mean",
                   "l2": "warm (--no-flush)" if flush is None else "flushed between steps (256 MiB memset, untimed)",
                   "parallelism": ("1 graph per GPU, dW summed over ranks by "
                                   + ("one kernel over peer-mapped buffers (NVLink, rank order)" if (world > 1 and not chain and par is not None)
                                      else "one coalesced NCCL all-reduce")
                                   + (" captured in the step's CUDA graph" if allreduce_in_graph else "")) if world > 1 else "single GPU",
                   "cuda_graph": bool(args.cuda_graph) and not chain},
        "rhs_evals_per_sec": K * world / (total_ms * 1e-3),
        "algorithmic_tflops": 3.0 * w.flops_fwd * world / (total_ms / K * 1e-3) / 1e12,
        "gpu_launches": launches,
        "gpu_launches_how": ("kernel nodes of the captured CUDA graph x replays (ngpde_cuda_graph_kernel_nodes)"
                             if graph_kernels is not None and args.cuda_graph and not chain
                             else "per-call table in ops.py"),
        "lib_sha256": lib_sha256(),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / K, "api": "layer(x, ps, st) + backward via the C ABI (torch.autograd bridge)"},
        "roofline": roofline,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        sample_kw = CPU_SAMPLE[args.workload]
        step, ws = cpu_reference_step_fn(args.workload, sample_kw)
        step()
        n, t0 = 0, time.perf_counter()
        while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 50):
            step()
            n += 1
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {
            "value": ws.n_edges * n / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{ws.name}: {ws.n_nodes} nodes / {ws.n_edges} edges "
                      f"({'the full workload' if not sample_kw else 'reduced: ' + str(sample_kw)}), {n} fwd+bwd steps in {dt:.1f} s "
                      "(oracle: torch-CPU restatement of the unfused reference algorithm; Julia is not installed)"}

    # ---- C1 is an ODE right-hand side on a graph that fits a thread-block cluster: the trajectory numbers ----
    if args.workload == "c1" and rank == 0 and world == 1:
        try:
            line["ode_trajectory"] = ode_trajectory_c1(w)
        except Exception as exc:  # noqa: BLE001
            line["ode_trajectory"] = {"error": repr(exc)}

    # ---- the north-star scaling configuration rides on the default line: C4, 1M nodes, node-partitioned, with parity ----
    if args.workload == "c3" and not args.no_strong_c4:
        del runner, w
        flush = None
        timed.flush = None
        torch.cuda.empty_cache()
        try:
            line["strong_c4"] = strong_c4(args, rank, world, dev, peaks)
        except Exception as exc:  # noqa: BLE001 -- never lose the headline line over the extra block
            import traceback
            line["strong_c4"] = {"error": repr(exc), "trace": traceback.format_exc()[-1500:]}
            if world > 1:
                raise

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
