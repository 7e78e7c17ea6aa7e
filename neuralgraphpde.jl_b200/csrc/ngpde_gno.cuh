// GNOConv, factored evaluation (layers.jl:509-547).
//
// When phi's last Dense layer is affine (no activation -- the reference's documented use, layers.jl:463-470) and the
// aggregation is + or mean, the per-edge kernel matrix never has to exist.  With z_e = phi_{1..L-1}(edge inputs) (K wide),
// h_e = x[src(e)] (gin wide) and the last layer  W_e = reshape(W3' z_e + b3, gout, gin):
//
//     sum_{e -> n} W_e h_e  =  B' vec(S_n),     S_n[j][i] = sum_{e -> n} za_e[j] * h_e[i],    za = [z; 1],
//
// where B = [W3; b3] viewed as a row-major [(K+1)*gin][gout] matrix is *exactly* the flat parameter segment of that layer
// (a Lux weight (gin*gout, K) column-major with output index o + gout*i, followed by its bias).  So
//
//   forward    edge kernel builds S [N][R] (R = (K+1)*gin; 4 KFLOP per edge instead of 524), one GEMM  mbar = S B;
//   backward   T = DM B'   (DM = dmbar ./ deg: the message cotangent is shared by all in-edges of a node),
//              per edge:  d h_e = T_n' za_e,   d z_e = T_n[0:K] h_e,    S rebuilt on the fly,   dB = S' DM.
//
// The contraction work drops from 2*gin*gout*K flop per EDGE to per NODE (x mean in-degree less), and all of it is plain
// dense GEMM.  Same sums in a different association order: float32 results agree with the per-edge evaluation to ~1e-6.
#pragma once
#include "ngpde_common.cuh"

namespace ngpde {

// C[M][N] (row stride ldc) = op(A) op(B) over K, FP32 FFMA, 128 x 64 tiles.
//   a_kmajor: A is stored [K][M] (row stride lda), else [M][K];   b_kmajor: B is stored [K][N] (ldb), else [N][K].
//   splits > 1: split-K, slice s writes C + s * M * ldc (caller reduces in fixed order).
//   deg_rowptr != nullptr: row m is divided by float(rowptr[m+1] - rowptr[m]) (rows with no in-edge give 0).
// Requires lda, ldb, ldc, and the contiguous extents to be multiples of 4 floats and 16-byte aligned bases.
int gno_gemm(const float* A, int lda, bool a_kmajor, const float* B, int ldb, bool b_kmajor, float* C, int ldc, int64_t M,
             int N, int64_t K, int splits, const int* deg_rowptr, cudaStream_t st);
// the two engines behind gno_gemm: FP32 FFMA (ngpde_gno.cu) and tcgen05 3xTF32 (ngpde_gno_tc.cu; taken when
// NGPDE_OPT_TENSOR_CORES is on and the strides allow it)
int gno_gemm_ffma(const float* A, int lda, bool a_kmajor, const float* B, int ldb, bool b_kmajor, float* C, int ldc,
                  int64_t M, int N, int64_t K, int splits, const int* deg_rowptr, cudaStream_t st);
bool gno_gemm_tc_supported(int lda, bool a_kmajor, int ldb, bool b_kmajor, int ldc, int N, int64_t K);
int gno_gemm_tc(const float* A, int lda, bool a_kmajor, const float* B, int ldb, bool b_kmajor, float* C, int ldc, int64_t M,
                int N, int64_t K, int splits, const int* deg_rowptr, cudaStream_t st);
bool tc_get_enabled();  // ngpde_tc.cu

// DM[n][:] = dmbar[n][:] / deg(n) for mean (true division, 0 for isolated nodes), plain copy for sum
int gno_dm_scale(const float* dmbar, const int* rowptr, int mean, int64_t N, int d, float* DM, cudaStream_t st);

}  // namespace ngpde
