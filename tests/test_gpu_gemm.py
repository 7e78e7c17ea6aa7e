"""Dense GEMM engines of the factored GNOConv evaluation (csrc/ngpde_gno.cu: FP32 FFMA; csrc/ngpde_gno_tc.cu: tcgen05 3xTF32
with a chunked FP32 flush of the TMEM accumulator) against a float64 product, through the C ABI (ngpde_debug_gemm), for the
three operand layouts the layer uses: mbar = S B, T = DM B', dB = S' DM (split-K).  Tolerance: 1e-5 of max|C| (the
per-layer bar of BASELINE.json's north_star); measured 2e-6."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("M,N,K,a_t,b_t,splits,deg", [
    (1000, 64, 4160, False, True, 1, True),     # mbar = (S B) ./ deg
    (1000, 4160, 64, False, False, 1, False),   # T = DM B'
    (4160, 64, 1000, True, True, 3, False),     # dB = S' DM, split-K
    (77, 12, 40, False, True, 1, True),         # ragged tiles
    (200, 20, 333, True, True, 2, False),
    (130, 36, 24, False, False, 1, False),
    (129, 64, 20000, False, True, 1, False),    # long K: the TMEM accumulator is flushed to FP32 registers every 256 k
])
def test_gemm_engines_against_float64(M, N, K, a_t, b_t, splits, deg):
    from gemm_check import case
    out = case("t", M, N, K, a_t, b_t, splits=splits, deg=deg)
    assert out["ffma"][0] <= TOL, out
    assert out["tcgen05"][0] <= TOL, out
