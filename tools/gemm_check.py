"""Checks the dense GEMM engines of the factored GNOConv (csrc/ngpde_gno.cu FFMA, csrc/ngpde_gno_tc.cu tcgen05 3xTF32)
against a float64 product, for the three operand layouts the layer uses, and times them at the C4 shapes.

    python tools/gemm_check.py [n_nodes]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ngpde
from ngpde import _lib


def gemm(A, a_t, B, b_t, M, N, K, splits=1, rowptr=None, engine=1):
    lib = _lib.load()
    C = torch.full((splits, M, N), float("nan"), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.ngpde_debug_gemm(A.data_ptr(), A.stride(0), int(a_t), B.data_ptr(), B.stride(0), int(b_t), C.data_ptr(), N,
                                    M, N, K, splits, rowptr.data_ptr() if rowptr is not None else None, engine, st))
    return C


def case(name, M, N, K, a_t, b_t, splits=1, deg=False, time_it=False):
    g = torch.Generator(device="cuda").manual_seed(M + 7 * N + 13 * K)
    A = torch.randn((K, M) if a_t else (M, K), device="cuda", generator=g)
    B = torch.randn((K, N) if b_t else (N, K), device="cuda", generator=g)
    Ad = (A.T if a_t else A).double()
    Bd = (B if b_t else B.T).double()
    ref = Ad @ Bd
    rowptr = None
    if deg:
        d = torch.randint(0, 5, (M,), device="cuda", generator=g)
        rowptr = torch.zeros(M + 1, dtype=torch.int32, device="cuda")
        rowptr[1:] = torch.cumsum(d, 0).int()
        ref = torch.where(d[:, None] > 0, ref / d[:, None].clamp(min=1).double(), torch.zeros_like(ref))
    out = {}
    for engine, ename in ((0, "ffma"), (1, "tcgen05")):
        C = gemm(A, a_t, B, b_t, M, N, K, splits, rowptr, engine).double().sum(0)
        err = float((C - ref).abs().max() / ref.abs().max())
        ms = None
        if time_it:
            for _ in range(2):
                gemm(A, a_t, B, b_t, M, N, K, splits, rowptr, engine)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                gemm(A, a_t, B, b_t, M, N, K, splits, rowptr, engine)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
        out[ename] = (err, ms)
    line = f"{name:34s} M={M} N={N} K={K} splits={splits}: " + "  ".join(
        f"{k} rel {v[0]:.2e}" + (f" {v[1]:.3f} ms {2.0 * M * N * K / (v[1] * 1e-3) / 1e12:.1f} TFLOP/s" if v[1] else "")
        for k, v in out.items())
    print(line, flush=True)
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    case("mbar = S B       (A [M][K], B [K][N])", 1000, 64, 4160, False, True, deg=True)
    case("T = DM B'        (A [M][K], B [N][K])", 1000, 4160, 64, False, False)
    case("dB = S' DM       (A [K][M], B [K][N])", 4160, 64, 1000, True, True, splits=3)
    case("ragged", 77, 12, 40, False, True)
    case("ragged T", 200, 20, 333, True, True, splits=2)
    case("ragged NT", 130, 36, 24, False, False)
    case("mbar = S B  @C4", n, 64, 4160, False, True, deg=True, time_it=True)
    case("T = DM B'   @C4", n, 4160, 64, False, False, time_it=True)
    case("dB = S' DM  @C4", 4160, 64, n, True, True, splits=32, time_it=True)
