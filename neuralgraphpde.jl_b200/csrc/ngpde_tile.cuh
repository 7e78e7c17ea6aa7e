// Tile machinery of the fused message-passing kernels (sm_100a, FP32 FFMA path).
//
// A CTA of 128 threads owns a tile of TE "columns" (edges in the edge phase, nodes in the node phase).
// Activations live in shared memory feature-major, Z[k][e] with row stride LD = TE + 4 floats, so that
//   * the layer GEMM  out[n][e] = sum_k W[k][n] * Z[k][e]  reads both operands as conflict-free float4s, and
//   * the weight-gradient GEMM  dW[k][n] = sum_e Z[k][e] * G[n][e]  reads rows with a 4-bank skew.
// Weights are streamed from global memory (L2-resident: a few hundred KB at most) through a double-buffered
// cp.async staging buffer of KC x NPASS floats; a Lux weight (out,in) column-major is exactly the [K][N]
// row-major operand this needs, so parameters are consumed in place from the flat ComponentArray.
#pragma once
#include "ngpde_common.cuh"

namespace ngpde {

constexpr int NT = 128;  // threads per CTA
constexpr int KC = 16;   // k-rows of W staged per pipeline step

template <int TE>
struct Cfg {
  static constexpr int RE = (TE >= 128) ? 8 : 4;  // columns per thread
  static constexpr int ETH = TE / RE;             // threads along the column axis
  static constexpr int NTH = NT / ETH;            // threads along the output-feature axis
  static constexpr int RN = 8;                    // output features per thread
  static constexpr int NPASS = NTH * RN;          // output features per pass
  static constexpr int LD = TE + 4;
  static constexpr int WS_FLOATS = 2 * KC * NPASS;
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(s), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

template <int TE>
__device__ __forceinline__ int e_of(int te, int i) {
  using C = Cfg<TE>;
  return (C::RE == 8 && i >= 4) ? (TE / 2 + te * 4 + (i - 4)) : (te * 4 + i);
}
template <int TE>
__device__ __forceinline__ int n_of(int tn, int j) {
  using C = Cfg<TE>;
  return (j >= 4) ? (C::NPASS / 2 + tn * 4 + (j - 4)) : (tn * 4 + j);
}

// Stage rows [k0, k0+KC) x columns [n0, n0+nvalid) of the row-major global matrix W (row stride ldw, K rows)
// into ws[KC][NPASS]; everything outside is zero-filled.
template <int TE>
__device__ __forceinline__ void load_w_chunk(float* ws, const float* __restrict__ W, int ldw, int K, int k0, int n0,
                                             int nvalid, int tid) {
  using C = Cfg<TE>;
  const float* base = W + n0;
  const bool vec = ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && ((ldw & 3) == 0);
  if (vec) {
    constexpr int Q = C::NPASS / 4;
    for (int i = tid; i < KC * Q; i += NT) {
      const int k = i / Q, c4 = (i - k * Q) * 4;
      const int gk = k0 + k, rem = nvalid - c4;
      const int bytes = (gk < K && rem > 0) ? (rem >= 4 ? 16 : rem * 4) : 0;
      const float* srcp = bytes ? (base + (size_t)gk * ldw + c4) : base;
      cp_async16(ws + k * C::NPASS + c4, srcp, bytes);
    }
  } else {
    for (int i = tid; i < KC * C::NPASS; i += NT) {
      const int k = i / C::NPASS, c = i - k * C::NPASS;
      const int gk = k0 + k;
      const int bytes = (gk < K && c < nvalid) ? 4 : 0;
      const float* srcp = bytes ? (base + (size_t)gk * ldw + c) : W;
      cp_async4(ws + k * C::NPASS + c, srcp, bytes);
    }
  }
}

// out-tile GEMM.  For pass p, colfn(p, n0, nvalid) names the columns of W it produces; after the last k-chunk
// of a pass, epi(p, acc) consumes the thread's RN x RE accumulator block:
//   acc[j][i] = sum_k W[k][n0 + n_of(tn, j)] * Zin[k][e_of(te, i)].
// Ends with a __syncthreads(), so whatever epi wrote to shared memory is visible on return.
template <int TE, class ColFn, class Epi>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ Zin, int K, const float* __restrict__ W, int ldw,
                                          int npass, ColFn colfn, float* ws, Epi epi) {
  using C = Cfg<TE>;
  const int tid = threadIdx.x;
  const int te = tid % C::ETH, tn = tid / C::ETH;
  const int nchunk = (K + KC - 1) / KC;
  const int total = npass * nchunk;
  if (total <= 0) return;
  float acc[C::RN][C::RE];
  {
    int n0, nv;
    colfn(0, n0, nv);
    load_w_chunk<TE>(ws, W, ldw, K, 0, n0, nv, tid);
    cp_commit();
  }
  int pass = 0, chunk = 0;
  for (int step = 0; step < total; ++step) {
    if (step + 1 < total) {
      int p2 = pass, c2 = chunk + 1;
      if (c2 == nchunk) { c2 = 0; ++p2; }
      int n0, nv;
      colfn(p2, n0, nv);
      load_w_chunk<TE>(ws + ((step + 1) & 1) * (KC * C::NPASS), W, ldw, K, c2 * KC, n0, nv, tid);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    if (chunk == 0) {
#pragma unroll
      for (int j = 0; j < C::RN; ++j)
#pragma unroll
        for (int i = 0; i < C::RE; ++i) acc[j][i] = 0.f;
    }
    const float* wsb = ws + (step & 1) * (KC * C::NPASS) + tn * 4;
    const float* zin = Zin + (size_t)chunk * KC * C::LD + te * 4;
    const int kmax = min(KC, K - chunk * KC);
    auto fma_step = [&](int k) {
      float a[C::RE], w[C::RN];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(zin + k * C::LD);
      if (C::RE == 8) *reinterpret_cast<float4*>(&a[C::RE - 4]) = *reinterpret_cast<const float4*>(zin + k * C::LD + TE / 2);
      *reinterpret_cast<float4*>(&w[0]) = *reinterpret_cast<const float4*>(wsb + k * C::NPASS);
      *reinterpret_cast<float4*>(&w[4]) = *reinterpret_cast<const float4*>(wsb + k * C::NPASS + C::NPASS / 2);
#pragma unroll
      for (int j = 0; j < C::RN; ++j)
#pragma unroll
        for (int i = 0; i < C::RE; ++i) acc[j][i] = fmaf(w[j], a[i], acc[j][i]);
    };
    if (kmax == KC) {
#pragma unroll
      for (int k = 0; k < KC; ++k) fma_step(k);
    } else {
      for (int k = 0; k < kmax; ++k) fma_step(k);
    }
    if (chunk == nchunk - 1) epi(pass, acc);
    __syncthreads();
    if (++chunk == nchunk) { chunk = 0; ++pass; }
  }
}

// Weight-gradient tile:  dW[k][n] += sum_{e < TE} A[k][e] * B[n][e]   (dW global, row stride ldw; A, B smem),
// and db[n] += sum_e B[n][e] when db != nullptr.  Every (k, n) is owned by one fixed thread, so the
// read-modify-write on the CTA-private partial buffer needs no synchronisation and is order-deterministic.
template <int TE>
__device__ __forceinline__ void tile_outer(float* __restrict__ dW, int ldw, float* __restrict__ db,
                                           const float* __restrict__ A, int K, const float* __restrict__ B, int N) {
  using C = Cfg<TE>;
  const int tid = threadIdx.x;
  const int kt = tid >> 4, nt = tid & 15;
  for (int kb = 0; kb < K; kb += 64) {
    for (int nb = 0; nb < N; nb += 64) {
      float acc[8][4];
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
      const float* ap[8];
      const float* bp[4];
#pragma unroll
      for (int r = 0; r < 8; ++r) ap[r] = A + (size_t)min(kb + kt + 8 * r, K - 1) * C::LD;
#pragma unroll
      for (int c = 0; c < 4; ++c) bp[c] = B + (size_t)min(nb + nt + 16 * c, N - 1) * C::LD;
      // rows this thread really owns in the block (a narrow first layer, K = 6, would otherwise cost a full 64-row block)
      const int nr = min(8, max(0, (K - kb - kt + 7) >> 3));
      if (nr == 8) {
#pragma unroll 2
        for (int e = 0; e < TE; e += 4) {
          float4 a[8], b[4];
#pragma unroll
          for (int r = 0; r < 8; ++r) a[r] = *reinterpret_cast<const float4*>(ap[r] + e);
#pragma unroll
          for (int c = 0; c < 4; ++c) b[c] = *reinterpret_cast<const float4*>(bp[c] + e);
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              acc[r][c] = fmaf(a[r].x, b[c].x, acc[r][c]);
              acc[r][c] = fmaf(a[r].y, b[c].y, acc[r][c]);
              acc[r][c] = fmaf(a[r].z, b[c].z, acc[r][c]);
              acc[r][c] = fmaf(a[r].w, b[c].w, acc[r][c]);
            }
        }
      } else {
        for (int e = 0; e < TE; e += 4) {
          float4 b[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) b[c] = *reinterpret_cast<const float4*>(bp[c] + e);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            if (r < nr) {
              const float4 a = *reinterpret_cast<const float4*>(ap[r] + e);
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                acc[r][c] = fmaf(a.x, b[c].x, acc[r][c]);
                acc[r][c] = fmaf(a.y, b[c].y, acc[r][c]);
                acc[r][c] = fmaf(a.z, b[c].z, acc[r][c]);
                acc[r][c] = fmaf(a.w, b[c].w, acc[r][c]);
              }
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int k = kb + kt + 8 * r;
        if (k < K) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int n = nb + nt + 16 * c;
            if (n < N) dW[(size_t)k * ldw + n] += acc[r][c];
          }
        }
      }
    }
  }
  if (db != nullptr) {
    for (int n = tid; n < N; n += NT) {
      float s = 0.f;
      const float* b = B + (size_t)n * C::LD;
      for (int e = 0; e < TE; ++e) s += b[e];
      db[n] += s;
    }
  }
}

}  // namespace ngpde
