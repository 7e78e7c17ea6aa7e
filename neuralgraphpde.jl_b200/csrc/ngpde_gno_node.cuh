// Per-destination part of the factored GNOConv evaluation (ngpde_gno.cuh) as a kernel of its own -- included by ngpde_conv.cu.
//
// The fused FFMA edge kernel (mp_bwd_kernel<64, edge>, contract == 2) runs phi's hidden layers AND the per-destination
// products on one 64-edge tile: 255 registers, 12.5 % of the warps resident, a barrier-separated chain per destination
// (C4: 82 of 116 ms at 1M nodes).  Here phi's hidden layers go through the tcgen05 GEMMs (ngpde_layered.cuh) and the
// per-destination products run in this kernel, ONE WARP PAIR PER DESTINATION NODE (the two warps share the node's staged
// operands and split the arithmetic: warp 0 takes dz and the lower half of S_n's rows, warp 1 dh and the upper half; they meet
// at a 64-thread named barrier), no block-level barrier anywhere:
//
//   forward    S_n[j][i] = sum over the in-edges e of n (ascending) of za_e[j] h_e[i]         za = [z; 1], h_e = x[src(e)]
//   backward   the same S_n (for dB = S' DM), and with T_n = DM_n B' staged once per node in the warp's shared memory:
//              dz_e[j] = sum_i T_n[j][i] h_e[i]  (j < K),      dh_e[i] = sum_j T_n[j][i] za_e[j]  (j < Ka)
//
// Edges are taken 16 at a time (ZT[j][e] and Hs[e][i] in shared memory; the arithmetic is instantiated for 4 / 8 / 12 / 16
// edges); lane l owns rows j = l, l + 32 of dz and columns i = l, l + 32 of dh and S, so every global access is a 128-byte
// row segment and every shared-memory access is conflict-free or a broadcast.  A node with more than 16 in-edges adds its
// later batches into the S_n it wrote itself.  Widths: K = gin in {32, 64}.
#pragma once
#include "ngpde_common.cuh"

namespace ngpde {
namespace gnonode {

constexpr int EB = 16;        // edges per batch
constexpr int LE = EB + 4;    // row stride of ZT[j][e]
constexpr int PAIRS = 8;      // warp pairs (= nodes in flight) per CTA

struct Args {
  const float* z;     // [E][K] last hidden activation of phi, CSR edge order
  const float* x;     // [N][ldx]
  int ldx;
  const int* src;     // [E] CSR order
  const int* rowptr;  // [N + 1]
  int N, Ka;          // Ka = K + 1 when phi's last layer has a bias, else K
  const float* T;     // [N][Ka * GIN]   (backward)
  float* S;           // [N][Ka * GIN]; null in the backward when the forward's S was kept (io.state)
  float* dz;          // [E][K]          (backward)
  float* desrc;       // [E][GIN]        (backward)
};

template <int KC, int GC, bool BWD>
struct Smem {
  static constexpr int K = 32 * KC, GIN = 32 * GC, LT = GIN + 4, LH = GIN + 4;
  static constexpr int ts = BWD ? (K + 1) * LT : 0;
  static constexpr int zt = (K + 1) * LE;
  static constexpr int hs = EB * LH;
  static constexpr int per_pair = ts + zt + hs;  // floats
  static constexpr int bytes = per_pair * PAIRS * 4;
};

template <int NE, int KC, int GC, bool BWD>
__device__ __forceinline__ void batch(const Args& a, const float* __restrict__ Ts, const float* __restrict__ ZT,
                                      const float* __restrict__ Hs, int n, int b0, int nb, bool first, int lane, int half) {
  constexpr int K = 32 * KC, GIN = 32 * GC, LT = GIN + 4, LH = GIN + 4;
  const int Ka = a.Ka;
  if (BWD && half == 0) {  // dz_e[j] = sum_i T[j][i] h_e[i], rows j = lane + 32 kc
    float acc[NE][KC];
#pragma unroll
    for (int e = 0; e < NE; ++e)
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) acc[e][kc] = 0.f;
#pragma unroll 2
    for (int i4 = 0; i4 < GIN / 4; ++i4) {
      float4 t[KC];
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) t[kc] = *reinterpret_cast<const float4*>(Ts + (lane + 32 * kc) * LT + 4 * i4);
#pragma unroll
      for (int e = 0; e < NE; ++e) {
        const float4 h = *reinterpret_cast<const float4*>(Hs + e * LH + 4 * i4);
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
          acc[e][kc] = fmaf(t[kc].x, h.x, acc[e][kc]);
          acc[e][kc] = fmaf(t[kc].y, h.y, acc[e][kc]);
          acc[e][kc] = fmaf(t[kc].z, h.z, acc[e][kc]);
          acc[e][kc] = fmaf(t[kc].w, h.w, acc[e][kc]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < NE; ++e)
      if (e < nb) {
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) a.dz[(size_t)(b0 + e) * K + lane + 32 * kc] = acc[e][kc];
      }
  }
  if (BWD && half == 1) {  // dh_e[i] = sum_j T[j][i] za_e[j], columns i = lane + 32 gc
    float acc[NE][GC];
#pragma unroll
    for (int e = 0; e < NE; ++e)
#pragma unroll
      for (int gc = 0; gc < GC; ++gc) acc[e][gc] = 0.f;
#pragma unroll 4
    for (int j = 0; j < Ka; ++j) {
      float t[GC];
#pragma unroll
      for (int gc = 0; gc < GC; ++gc) t[gc] = Ts[j * LT + lane + 32 * gc];
#pragma unroll
      for (int e4 = 0; e4 < NE / 4; ++e4) {
        const float4 za = *reinterpret_cast<const float4*>(ZT + j * LE + 4 * e4);
#pragma unroll
        for (int gc = 0; gc < GC; ++gc) {
          acc[4 * e4 + 0][gc] = fmaf(t[gc], za.x, acc[4 * e4 + 0][gc]);
          acc[4 * e4 + 1][gc] = fmaf(t[gc], za.y, acc[4 * e4 + 1][gc]);
          acc[4 * e4 + 2][gc] = fmaf(t[gc], za.z, acc[4 * e4 + 2][gc]);
          acc[4 * e4 + 3][gc] = fmaf(t[gc], za.w, acc[4 * e4 + 3][gc]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < NE; ++e)
      if (e < nb) {
#pragma unroll
        for (int gc = 0; gc < GC; ++gc) a.desrc[(size_t)(b0 + e) * GIN + lane + 32 * gc] = acc[e][gc];
      }
  }
  if (a.S != nullptr) {  // S_n[j][i] (+)= sum_e za_e[j] h_e[i], ascending e; columns i = lane + 32 gc; this warp's half of the rows j
    float h[NE][GC];
#pragma unroll
    for (int e = 0; e < NE; ++e)
#pragma unroll
      for (int gc = 0; gc < GC; ++gc) h[e][gc] = Hs[e * LH + lane + 32 * gc];
    float* __restrict__ Sn = a.S + (size_t)n * Ka * GIN;
    const int jm = (Ka + 1) >> 1;
    const int j0 = half ? jm : 0, j1 = half ? Ka : jm;
#pragma unroll 4
    for (int j = j0; j < j1; ++j) {
      float s[GC];
#pragma unroll
      for (int gc = 0; gc < GC; ++gc) s[gc] = first ? 0.f : Sn[j * GIN + lane + 32 * gc];
#pragma unroll
      for (int e4 = 0; e4 < NE / 4; ++e4) {
        const float4 za = *reinterpret_cast<const float4*>(ZT + j * LE + 4 * e4);
#pragma unroll
        for (int gc = 0; gc < GC; ++gc) {
          s[gc] = fmaf(za.x, h[4 * e4 + 0][gc], s[gc]);
          s[gc] = fmaf(za.y, h[4 * e4 + 1][gc], s[gc]);
          s[gc] = fmaf(za.z, h[4 * e4 + 2][gc], s[gc]);
          s[gc] = fmaf(za.w, h[4 * e4 + 3][gc], s[gc]);
        }
      }
#pragma unroll
      for (int gc = 0; gc < GC; ++gc) Sn[j * GIN + lane + 32 * gc] = s[gc];
    }
  }
}

__device__ __forceinline__ void pair_sync(int pair) { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int KC, int GC, bool BWD>
__global__ void __launch_bounds__(64 * PAIRS, BWD ? 1 : 2) gno_node_kernel(const Args a) {
  using SM = Smem<KC, GC, BWD>;
  constexpr int K = SM::K, GIN = SM::GIN, LT = SM::LT, LH = SM::LH;
  extern __shared__ __align__(16) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pair = warp >> 1, half = warp & 1;
  float* Ts = smem + pair * SM::per_pair;
  float* ZT = Ts + SM::ts;
  float* Hs = ZT + SM::zt;
  const int Ka = a.Ka;
  const int R = Ka * GIN;
  for (int n = blockIdx.x * PAIRS + pair; n < a.N; n += gridDim.x * PAIRS) {
    const int r0 = a.rowptr[n], r1 = a.rowptr[n + 1];
    if (!BWD && half == 0) {  // the pair's NEXT node: the z rows of its first batch are pulled into L2 while this node computes
      const int nn = n + gridDim.x * PAIRS;
      if (nn < a.N) {
        const int q0 = a.rowptr[nn], q1 = a.rowptr[nn + 1];
        const int lines = (min(q1 - q0, EB) * K * 4 + 127) / 128;
        if (lane < lines) prefetch_l2(reinterpret_cast<const char*>(a.z + (size_t)q0 * K) + (size_t)lane * 128);
      }
    }
    if (r0 == r1) {  // isolated destination: S_n = 0 (mbar = 0, nothing for dB)
      if (a.S != nullptr) {
        float4* Sn = reinterpret_cast<float4*>(a.S + (size_t)n * R);
        for (int i = half * 32 + lane; i < R / 4; i += 64) Sn[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      continue;
    }
    if (BWD) {
      const float4* __restrict__ Tn = reinterpret_cast<const float4*>(a.T + (size_t)n * R);
#pragma unroll 8
      for (int i = half * 32 + lane; i < R / 4; i += 64) {
        const float4 v = __ldg(Tn + i);
        const int j = (4 * i) / GIN, c = (4 * i) - j * GIN;
        *reinterpret_cast<float4*>(Ts + j * LT + c) = v;
      }
    }
    for (int b0 = r0; b0 < r1; b0 += EB) {
      const int nb = min(EB, r1 - b0);
      if (half == 0) {  // ZT[j][e] = z_e[j], the bias row of za (read only when Ka == K + 1)
#pragma unroll 8
        for (int e = 0; e < EB; ++e) {
          const bool on = e < nb;
#pragma unroll
          for (int kc = 0; kc < KC; ++kc)
            ZT[(lane + 32 * kc) * LE + e] = on ? a.z[(size_t)(b0 + e) * K + lane + 32 * kc] : 0.f;
        }
        if (lane < EB) ZT[K * LE + lane] = lane < nb ? 1.f : 0.f;
      } else {          // Hs[e][i] = x[src(e)][i]
        const int my_src = lane < nb ? a.src[b0 + lane] : 0;
#pragma unroll 8
        for (int e = 0; e < EB; ++e) {
          const bool on = e < nb;
          const int s = __shfl_sync(0xffffffffu, my_src, e);
#pragma unroll
          for (int gc = 0; gc < GC; ++gc) Hs[e * LH + lane + 32 * gc] = on ? __ldg(a.x + (size_t)s * a.ldx + lane + 32 * gc) : 0.f;
        }
      }
      pair_sync(pair);
      const bool first = b0 == r0;
      switch ((nb + 3) >> 2) {
        case 1: batch<4, KC, GC, BWD>(a, Ts, ZT, Hs, n, b0, nb, first, lane, half); break;
        case 2: batch<8, KC, GC, BWD>(a, Ts, ZT, Hs, n, b0, nb, first, lane, half); break;
        case 3: batch<12, KC, GC, BWD>(a, Ts, ZT, Hs, n, b0, nb, first, lane, half); break;
        default: batch<16, KC, GC, BWD>(a, Ts, ZT, Hs, n, b0, nb, first, lane, half); break;
      }
      pair_sync(pair);
    }
  }
}

inline bool supported(int K, int gin) { return K == gin && (K == 32 || K == 64); }

template <int KC, int GC, bool BWD>
int launch_one(const Args& a, int num_sms, cudaStream_t st) {
  using SM = Smem<KC, GC, BWD>;
  static bool configured = false;
  if (!configured) {
    NGPDE_CUDA_TRY(cudaFuncSetAttribute(gno_node_kernel<KC, GC, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::bytes));
    configured = true;
  }
  const int per_sm = BWD ? 1 : std::max(1, std::min(2, (220 * 1024) / SM::bytes));
  const int grid = std::max(1, std::min(num_sms * per_sm, (a.N + PAIRS - 1) / PAIRS));
  gno_node_kernel<KC, GC, BWD><<<grid, 64 * PAIRS, SM::bytes, st>>>(a);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

inline int launch(const Args& a, int K, bool backward, int num_sms, cudaStream_t st) {
  if (K == 64) return backward ? launch_one<2, 2, true>(a, num_sms, st) : launch_one<2, 2, false>(a, num_sms, st);
  return backward ? launch_one<1, 1, true>(a, num_sms, st) : launch_one<1, 1, false>(a, num_sms, st);
}

}  // namespace gnonode
}  // namespace ngpde
