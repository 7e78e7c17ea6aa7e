// Dense FP32 GEMMs of the factored GNOConv evaluation (see ngpde_gno.cuh) and the cotangent pre-scale.
//
// 128 x 64 output tile per CTA of 128 threads, 8 x 8 outputs per thread, BK = 16, operands staged k-major in shared
// memory (register-staged double buffering: the next k-slab's global loads are in flight during the FMAs).  A warp's
// lanes are 8 wide along N and 4 along M, so C rows are written as full 128-byte segments and both operand reads
// are single-wavefront float4 broadcasts.
#include "ngpde_gno.cuh"

namespace ngpde {
namespace {

constexpr int BM = 128, BN = 64, BK = 16, GT = 128;
constexpr int LDA_S = BM + 4, LDB_S = BN + 4;

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  const int* deg_rowptr;
  long long M, K;
  int N, lda, ldb, ldc;
  long long k_per_split;
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Loads one BK-slab of an operand tile into registers.  ROWS = tile extent along the non-k axis (BM or BN).
//   KMAJOR: source is [K][rows] -> float4 along rows;  else source is [rows][K] -> float4 along k.
template <int ROWS, bool KMAJOR>
__device__ __forceinline__ void load_slab(float4 (&r)[ROWS * BK / 4 / GT], const float* __restrict__ P, int ld,
                                          long long row0, long long nrows, long long k0, long long kend, int tid) {
  constexpr int NV = ROWS * BK / 4 / GT;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int idx = tid + v * GT;
    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
    if (KMAJOR) {
      constexpr int Q = ROWS / 4;
      const int k = idx / Q, c4 = (idx - k * Q) * 4;
      const long long gk = k0 + k, gr = row0 + c4;
      if (gk < kend && gr < nrows) val = ldg4(P + (size_t)gk * ld + gr);  // nrows % 4 == 0
    } else {
      const int row = idx / (BK / 4), kq = (idx - row * (BK / 4)) * 4;
      const long long gr = row0 + row, gk = k0 + kq;
      if (gr < nrows && gk < kend) val = ldg4(P + (size_t)gr * ld + gk);  // kend % 4 == 0
    }
    r[v] = val;
  }
}

template <int ROWS, bool KMAJOR>
__device__ __forceinline__ void store_slab(const float4 (&r)[ROWS * BK / 4 / GT], float* __restrict__ S, int lds, int tid) {
  constexpr int NV = ROWS * BK / 4 / GT;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int idx = tid + v * GT;
    if (KMAJOR) {
      constexpr int Q = ROWS / 4;
      const int k = idx / Q, c4 = (idx - k * Q) * 4;
      *reinterpret_cast<float4*>(S + k * lds + c4) = r[v];
    } else {
      const int row = idx / (BK / 4), kq = (idx - row * (BK / 4)) * 4;
      S[(kq + 0) * lds + row] = r[v].x;
      S[(kq + 1) * lds + row] = r[v].y;
      S[(kq + 2) * lds + row] = r[v].z;
      S[(kq + 3) * lds + row] = r[v].w;
    }
  }
}

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(GT) gno_gemm_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[2][BK * LDA_S];
  __shared__ __align__(16) float Bs[2][BK * LDB_S];
  const int tid = threadIdx.x;
  const int tn = tid & 7, tm = tid >> 3;  // 8 lanes along N, 16 along M
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const long long kbeg = (long long)blockIdx.z * g.k_per_split;
  const long long kend = min(g.K, kbeg + g.k_per_split);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[BM * BK / 4 / GT], rb[BN * BK / 4 / GT];
  const long long nslab = (kend > kbeg) ? (kend - kbeg + BK - 1) / BK : 0;
  if (nslab > 0) {
    load_slab<BM, A_KMAJOR>(ra, g.A, g.lda, m0, g.M, kbeg, kend, tid);
    load_slab<BN, B_KMAJOR>(rb, g.B, g.ldb, n0, g.N, kbeg, kend, tid);
    store_slab<BM, A_KMAJOR>(ra, As[0], LDA_S, tid);
    store_slab<BN, B_KMAJOR>(rb, Bs[0], LDB_S, tid);
  }
  __syncthreads();
  for (long long s = 0; s < nslab; ++s) {
    const int cur = (int)(s & 1);
    if (s + 1 < nslab) {
      load_slab<BM, A_KMAJOR>(ra, g.A, g.lda, m0, g.M, kbeg + (s + 1) * BK, kend, tid);
      load_slab<BN, B_KMAJOR>(rb, g.B, g.ldb, n0, g.N, kbeg + (s + 1) * BK, kend, tid);
    }
    const float* as = As[cur] + tm * 4;
    const float* bs = Bs[cur] + tn * 4;
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(as + k * LDA_S);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(as + k * LDA_S + BM / 2);
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(bs + k * LDB_S);
      *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(bs + k * LDB_S + BN / 2);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (s + 1 < nslab) {
      store_slab<BM, A_KMAJOR>(ra, As[cur ^ 1], LDA_S, tid);
      store_slab<BN, B_KMAJOR>(rb, Bs[cur ^ 1], LDB_S, tid);
    }
    __syncthreads();
  }

  float* C = g.C + (size_t)blockIdx.z * (size_t)g.M * g.ldc;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long m = m0 + (i < 4 ? tm * 4 + i : BM / 2 + tm * 4 + (i - 4));
    if (m >= g.M) continue;
    float inv_den = 0.f;
    bool scale = false, zero = false;
    if (g.deg_rowptr != nullptr) {
      const int deg = g.deg_rowptr[m + 1] - g.deg_rowptr[m];
      scale = true;
      zero = deg == 0;
      inv_den = (float)deg;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = n0 + (h == 0 ? tn * 4 : BN / 2 + tn * 4);
      if (n >= g.N) continue;  // N % 4 == 0
      float4 v = make_float4(acc[i][4 * h + 0], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
      if (scale) {
        if (zero) {
          v = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
          v.x = __fdiv_rn(v.x, inv_den); v.y = __fdiv_rn(v.y, inv_den);
          v.z = __fdiv_rn(v.z, inv_den); v.w = __fdiv_rn(v.w, inv_den);
        }
      }
      *reinterpret_cast<float4*>(C + (size_t)m * g.ldc + n) = v;
    }
  }
}

__global__ void gno_dm_scale_kernel(const float* __restrict__ dmbar, const int* __restrict__ rowptr, int mean,
                                    long long total, int d, float* __restrict__ DM) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const long long n = idx / d;
  float v = dmbar[idx];
  const int deg = rowptr[n + 1] - rowptr[n];
  if (deg == 0) v = 0.f;
  else if (mean) v = __fdiv_rn(v, (float)deg);
  DM[idx] = v;
}

}  // namespace

int gno_gemm(const float* A, int lda, bool a_kmajor, const float* B, int ldb, bool b_kmajor, float* C, int ldc, int64_t M,
             int N, int64_t K, int splits, const int* deg_rowptr, cudaStream_t st) {
  // short-K products are bound by the output write and per-tile overheads, where the generic tensor-core kernel has no
  // edge over the FFMA one (measured at C4: 4.1 vs 3.9 ms); T = DM B' has its own variant that walks the n-tiles
  const bool nloop = !a_kmajor && !b_kmajor && K <= 64 && splits <= 1 && deg_rowptr == nullptr && N > 64;  // T = DM B'
  if (tc_get_enabled() && (K > 128 || nloop) && gno_gemm_tc_supported(lda, a_kmajor, ldb, b_kmajor, ldc, N, K))
    return gno_gemm_tc(A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, M, N, K, splits, deg_rowptr, st);
  return gno_gemm_ffma(A, lda, a_kmajor, B, ldb, b_kmajor, C, ldc, M, N, K, splits, deg_rowptr, st);
}

int gno_gemm_ffma(const float* A, int lda, bool a_kmajor, const float* B, int ldb, bool b_kmajor, float* C, int ldc,
                  int64_t M, int N, int64_t K, int splits, const int* deg_rowptr, cudaStream_t st) {
  if (M <= 0 || N <= 0) return NGPDE_OK;
  NGPDE_REQUIRE((lda & 3) == 0 && (ldb & 3) == 0 && (ldc & 3) == 0 && (N & 3) == 0, "gno_gemm: strides must be multiples of 4");
  NGPDE_REQUIRE(a_kmajor ? (M & 3) == 0 : (K & 3) == 0, "gno_gemm: contiguous extent of A must be a multiple of 4");
  NGPDE_REQUIRE(b_kmajor || (K & 3) == 0, "gno_gemm: contiguous extent of B must be a multiple of 4");
  GemmArgs g;
  g.A = A; g.B = B; g.C = C; g.deg_rowptr = deg_rowptr;
  g.M = M; g.K = K; g.N = N; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  splits = splits < 1 ? 1 : splits;
  long long kps = (K + splits - 1) / splits;
  kps = (kps + BK - 1) / BK * BK;  // slabs stay 16-aligned inside every slice
  g.k_per_split = kps > 0 ? kps : BK;
  dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)((N + BN - 1) / BN), (unsigned)splits);
  if (a_kmajor && b_kmajor) gno_gemm_kernel<true, true><<<grid, GT, 0, st>>>(g);
  else if (!a_kmajor && b_kmajor) gno_gemm_kernel<false, true><<<grid, GT, 0, st>>>(g);
  else if (!a_kmajor && !b_kmajor) gno_gemm_kernel<false, false><<<grid, GT, 0, st>>>(g);
  else gno_gemm_kernel<true, false><<<grid, GT, 0, st>>>(g);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

int gno_dm_scale(const float* dmbar, const int* rowptr, int mean, int64_t N, int d, float* DM, cudaStream_t st) {
  const long long total = (long long)N * d;
  if (total == 0) return NGPDE_OK;
  gno_dm_scale_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dmbar, rowptr, mean, total, d, DM);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

}  // namespace ngpde

extern "C" int ngpde_debug_gemm(const float* A, int32_t lda, int32_t a_kmajor, const float* B, int32_t ldb, int32_t b_kmajor,
                                float* C, int32_t ldc, int64_t M, int32_t N, int64_t K, int32_t splits,
                                const int32_t* deg_rowptr, int32_t engine, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (engine == 1)
    return ngpde::gno_gemm_tc(A, lda, a_kmajor != 0, B, ldb, b_kmajor != 0, C, ldc, M, N, K, splits, deg_rowptr, st);
  return ngpde::gno_gemm_ffma(A, lda, a_kmajor != 0, B, ldb, b_kmajor != 0, C, ldc, M, N, K, splits, deg_rowptr, st);
}
