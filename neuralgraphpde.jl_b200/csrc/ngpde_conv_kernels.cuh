// Fused message-passing kernels: gather -> MLP (phi) -> ordered aggregate (edge phase), and the node-update MLP
// (gamma / psi / linear) as the same engine over an identity topology (node phase).  Forward and backward.
//
// Replaces, for ExplicitEdgeConv / VMHConv / MPPDEConv / GNOConv (/root/reference/src/layers.jl:98-112, 312-332,
// 390-422, 509-547), the dependency chain  NNlib.gather -> vcat -> Lux.Dense (sgemm + broadcast) ...
// -> NNlib.scatter  that GraphNeuralNetworks.propagate executes.
#pragma once
#include "ngpde_conv.cuh"
#include "ngpde_gno_tile.cuh"

namespace ngpde {

__device__ __forceinline__ float aggr_identity(int aggr) {
  return aggr == NGPDE_AGGR_MAX ? -INFINITY : (aggr == NGPDE_AGGR_MIN ? INFINITY : (aggr == NGPDE_AGGR_PROD ? 1.f : 0.f));
}

template <int TE>
__device__ __forceinline__ void load_ids(const TileGraph& tg, bool node, int k0, int ne, int* s_src, int* s_dst,
                                         int* s_perm) {
  const int tid = threadIdx.x;
  if (tid < TE) {
    int s = 0, d = 0, p = 0;
    if (tid < ne) {
      if (node) {
        s = d = p = k0 + tid;
      } else {
        s = tg.src[k0 + tid];
        d = tg.dst[k0 + tid];
        p = tg.perm[k0 + tid];
      }
    }
    s_src[tid] = s;
    s_dst[tid] = d;
    s_perm[tid] = p;
  }
}

// Z0[seg.row + f][e] for every segment; columns e >= ne are zero-filled.
template <int TE, class Args>
__device__ __forceinline__ void gather_tile(float* __restrict__ Z, const Args& a, int ne, const int* s_src,
                                            const int* s_dst, const int* s_perm) {
  using C = Cfg<TE>;
  const int tid = threadIdx.x;
  for (int si = 0; si < a.n_segs; ++si) {
    const Seg sg = a.segs[si];
    const float* __restrict__ A = a.arr[sg.arr] + sg.col;
    const int ld = a.ld[sg.arr];
    const int total = sg.width * TE;
    const bool wide = sg.width >= 16;
    for (int item = tid; item < total; item += NT) {
      int f, e;
      if (wide) {
        e = item / sg.width;
        f = item - e * sg.width;
      } else {
        f = item / TE;
        e = item - f * TE;
      }
      float v = 0.f;
      if (e < ne) {
        switch (sg.kind) {
          case SEG_DST: v = A[(size_t)s_dst[e] * ld + f]; break;
          case SEG_SRC: v = A[(size_t)s_src[e] * ld + f]; break;
          case SEG_SMD: v = A[(size_t)s_src[e] * ld + f] - A[(size_t)s_dst[e] * ld + f]; break;
          case SEG_DMS: v = A[(size_t)s_dst[e] * ld + f] - A[(size_t)s_src[e] * ld + f]; break;
          case SEG_EDGE: v = A[(size_t)s_perm[e] * ld + f]; break;
          default: v = A[(size_t)(s_perm[e] / a.tg.gdiv) * ld + f]; break;  // SEG_GRAPH
        }
      }
      Z[(sg.row + f) * C::LD + e] = v;
    }
  }
}

// One Dense layer on the tile: Zout[n][e] = act(sum_k W[k][n] Zin[k][e] + b[n] (+ addend[col e][n])).
template <int TE>
__device__ __forceinline__ void dense_tile(const float* __restrict__ Zin, float* __restrict__ Zout, int K, int N,
                                           const float* __restrict__ W, const float* __restrict__ bias, int act,
                                           float* ws, const float* __restrict__ addend, int col0, int ne) {
  using C = Cfg<TE>;
  const int tid = threadIdx.x;
  const int te = tid % C::ETH, tn = tid / C::ETH;
  const int npass = (N + C::NPASS - 1) / C::NPASS;
  auto colfn = [&](int p, int& n0, int& nv) {
    n0 = p * C::NPASS;
    nv = min(C::NPASS, N - n0);
  };
  auto epi = [&](int p, float (&acc)[C::RN][C::RE]) {
#pragma unroll
    for (int j = 0; j < C::RN; ++j) {
      const int n = p * C::NPASS + n_of<TE>(tn, j);
      if (n < N) {
        const float b = bias ? bias[n] : 0.f;
        float v[C::RE];
#pragma unroll
        for (int i = 0; i < C::RE; ++i) {
          float pre = acc[j][i] + b;
          if (addend != nullptr) {
            const int e = e_of<TE>(te, i);
            if (e < ne) pre = (acc[j][i] + addend[(size_t)(col0 + e) * N + n]) + b;
          }
          v[i] = act_fwd(act, pre);
        }
        float* o = Zout + (size_t)n * C::LD + te * 4;
        *reinterpret_cast<float4*>(o) = *reinterpret_cast<float4*>(&v[0]);
        if (C::RE == 8) *reinterpret_cast<float4*>(o + TE / 2) = *reinterpret_cast<float4*>(&v[C::RE - 4]);
      }
    }
  };
  tile_gemm<TE>(Zin, K, W, N, npass, colfn, ws, epi);
}

// Sequential per-destination reduction of the message tile M[dm][LD] in ascending CSR (= original edge) order,
// carried across tiles of the same row through `mbar` itself.  No atomics: one thread owns (row, channel).
template <int TE>
__device__ __forceinline__ void aggregate_tile(const float* __restrict__ M, int dm, int aggr,
                                               const int* __restrict__ rowptr, int n0, int n1, int k0, int ne,
                                               float* __restrict__ mbar) {
  using C = Cfg<TE>;
  const int total = (n1 - n0) * dm;
  for (int item = threadIdx.x; item < total; item += NT) {
    const int jj = item / dm, c = item - jj * dm;
    const int j = n0 + jj;
    const int r0 = rowptr[j], r1 = rowptr[j + 1];
    const int lo = max(r0, k0), hi = min(r1, k0 + ne);
    if (lo >= hi) continue;
    float acc = (lo == r0) ? aggr_identity(aggr) : mbar[(size_t)j * dm + c];
    const float* m = M + (size_t)c * C::LD - k0;
    if (aggr == NGPDE_AGGR_MAX) {
      for (int e = lo; e < hi; ++e) acc = fmaxf(acc, m[e]);
    } else if (aggr == NGPDE_AGGR_MIN) {
      for (int e = lo; e < hi; ++e) acc = fminf(acc, m[e]);
    } else if (aggr == NGPDE_AGGR_PROD) {
      for (int e = lo; e < hi; ++e) acc = __fmul_rn(acc, m[e]);
    } else {
      for (int e = lo; e < hi; ++e) acc = __fadd_rn(acc, m[e]);
      if (aggr == NGPDE_AGGR_MEAN && hi == r1) acc = __fdiv_rn(acc, (float)(r1 - r0));
    }
    mbar[(size_t)j * dm + c] = acc;
  }
}

template <int TE, bool NODE>
__global__ void __launch_bounds__(NT) mp_fwd_kernel(const FwdArgs a) {
  using C = Cfg<TE>;
  extern __shared__ __align__(16) float smem[];
  int* s_src = reinterpret_cast<int*>(smem);
  int* s_dst = s_src + TE;
  int* s_perm = s_dst + TE;
  float* base = smem + 3 * TE;
  float* bufA = base + a.offA;
  float* bufB = base + a.offB;
  float* ws = base + a.offW;
  float* H = base + a.offH;
  const int tid = threadIdx.x;
  const MlpDev& mlp = a.mlp;
  const int Lp = a.contract ? mlp.L - 1 : mlp.L;

  for (int unit = blockIdx.x; unit < a.tg.n_units; unit += gridDim.x) {
    int n0, n1, kbeg, kend;
    if (NODE) {
      n0 = unit * TE;
      n1 = min(a.tg.N, n0 + TE);
      kbeg = n0;
      kend = n1;
    } else {
      n0 = a.tg.unit_ptr[unit];
      n1 = a.tg.unit_ptr[unit + 1];
      kbeg = a.tg.rowptr[n0];
      kend = a.tg.rowptr[n1];
      // isolated destinations get the identity of the reduction (0 for sum and mean)
      const float ident = aggr_identity(a.aggr);
      if (a.contract == 2) {
        gno_zero_isolated(a.tg.rowptr, n0, n1, (size_t)a.gno_Ka * a.gin, a.gno_S);
      } else if (a.msg_out == nullptr) {
        for (int item = tid; item < (n1 - n0) * a.dout; item += NT) {
          const int jj = item / a.dout;
          if (a.tg.rowptr[n0 + jj] == a.tg.rowptr[n0 + jj + 1]) a.out[(size_t)n0 * a.dout + item] = ident;
        }
      }
    }
    for (int k0 = kbeg; k0 < kend; k0 += TE) {
      const int ne = min(TE, kend - k0);
      load_ids<TE>(a.tg, NODE, k0, ne, s_src, s_dst, s_perm);
      __syncthreads();
      gather_tile<TE>(bufA, a, ne, s_src, s_dst, s_perm);
      if (a.contract == 2) {
        gno_gather_h<TE>(a.arr[ARR_X], a.ld[ARR_X], a.gin, s_src, ne, H, a.gin + 4);
      } else if (a.contract) {
        const float* __restrict__ X = a.arr[ARR_X];
        const int ldx = a.ld[ARR_X];
        for (int item = tid; item < a.gin * TE; item += NT) {
          int e, f;
          if (a.gin >= 16) { e = item / a.gin; f = item - e * a.gin; } else { f = item / TE; e = item - f * TE; }
          H[f * C::LD + e] = (e < ne) ? X[(size_t)s_src[e] * ldx + f] : 0.f;
        }
      }
      __syncthreads();
      float* cur = bufA;
      float* nxt = bufB;
      for (int l = 0; l < Lp; ++l) {
        const float* bias = mlp.b_off[l] >= 0 ? a.params + mlp.b_off[l] : nullptr;
        const float* add = (NODE && l == mlp.L - 1) ? a.addend : nullptr;
        dense_tile<TE>(cur, nxt, mlp.dims[l], mlp.dims[l + 1], a.params + mlp.w_off[l], bias, mlp.act[l], ws, add,
                       k0, ne);
        float* t = cur; cur = nxt; nxt = t;
      }
      if (a.contract == 2) {
        // GNOConv, factored: S_n += [z_e; 1] h_e' for the rows of this tile; mbar = S B follows as one GEMM
        const int K = mlp.dims[mlp.L - 1];
        float* Zt = base + a.offZt;
        const int ldz = gno_ldz(a.gno_Ka);
        gno_transpose_z<TE>(cur, K, a.gno_Ka > K, ne, Zt, ldz);
        __syncthreads();
        gno_outer_rows<TE>(Zt, ldz, H, a.gin + 4, a.gno_Ka, a.gin, a.tg.rowptr, n0, n1, k0, ne, a.gno_S);
        __syncthreads();
        continue;
      }
      if (a.contract) {
        // GNOConv: m[o][e] = sum_i act(phi_L(z))[o + gout*i][e] * h_src[i][e]; the (in*out) kernel matrix of an
        // edge is produced 64 entries at a time in registers and consumed immediately (layers.jl:523-530).
        const int l = mlp.L - 1;
        const int K = mlp.dims[l], NL = mlp.dims[l + 1];
        const float* W = a.params + mlp.w_off[l];
        const float* bias = mlp.b_off[l] >= 0 ? a.params + mlp.b_off[l] : nullptr;
        const int act = mlp.act[l];
        const int te = tid % C::ETH, tn = tid / C::ETH;
        for (int ob = 0; ob * C::NPASS < a.gout; ++ob) {
          const int ncols = min(C::NPASS, a.gout - ob * C::NPASS);
          float macc[C::RN][C::RE];
#pragma unroll
          for (int j = 0; j < C::RN; ++j)
#pragma unroll
            for (int i = 0; i < C::RE; ++i) macc[j][i] = 0.f;
          auto colfn = [&](int p, int& c0, int& nv) {
            c0 = a.gout * p + ob * C::NPASS;
            nv = ncols;
          };
          auto epi = [&](int p, float (&acc)[C::RN][C::RE]) {
            float h[C::RE];
#pragma unroll
            for (int i = 0; i < C::RE; ++i) h[i] = H[p * C::LD + e_of<TE>(te, i)];
#pragma unroll
            for (int j = 0; j < C::RN; ++j) {
              const int ol = n_of<TE>(tn, j);
              if (ol < ncols) {
                const float b = bias ? bias[a.gout * p + ob * C::NPASS + ol] : 0.f;
#pragma unroll
                for (int i = 0; i < C::RE; ++i) macc[j][i] = fmaf(act_fwd(act, acc[j][i] + b), h[i], macc[j][i]);
              }
            }
          };
          tile_gemm<TE>(cur, K, W, NL, a.gin, colfn, ws, epi);
#pragma unroll
          for (int j = 0; j < C::RN; ++j) {
            const int ol = n_of<TE>(tn, j);
            if (ol < ncols) {
              float* o = nxt + (size_t)(ob * C::NPASS + ol) * C::LD + te * 4;
              *reinterpret_cast<float4*>(o) = *reinterpret_cast<float4*>(&macc[j][0]);
              if (C::RE == 8) *reinterpret_cast<float4*>(o + TE / 2) = *reinterpret_cast<float4*>(&macc[j][C::RE - 4]);
            }
          }
        }
        __syncthreads();
        float* t = cur; cur = nxt; nxt = t;
      }
      if (NODE) {
        const int d = a.dout;
        for (int item = tid; item < ne * d; item += NT) {
          const int e = item / d, c = item - e * d;
          a.out[(size_t)(k0 + e) * d + c] = cur[c * C::LD + e];
        }
      } else if (a.msg_out != nullptr) {
        const int d = a.dout;
        for (int item = tid; item < ne * d; item += NT) {
          const int e = item / d, c = item - e * d;
          a.msg_out[(size_t)(k0 + e) * d + c] = cur[c * C::LD + e];
        }
      } else {
        aggregate_tile<TE>(cur, a.dout, a.aggr, a.tg.rowptr, n0, n1, k0, ne, a.out);
      }
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------


template <int TE, bool NODE>
__global__ void __launch_bounds__(NT) mp_bwd_kernel(const BwdArgs a) {
  using C = Cfg<TE>;
  extern __shared__ __align__(16) float smem[];
  int* s_src = reinterpret_cast<int*>(smem);
  int* s_dst = s_src + TE;
  int* s_perm = s_dst + TE;
  float* base = smem + 3 * TE;
  float* ws = base + a.offW;
  const int tid = threadIdx.x;
  const int te = tid % C::ETH, tn = tid / C::ETH;
  const MlpDev& mlp = a.mlp;
  const int L = mlp.L;
  const int Lp = a.contract ? L - 1 : L;
  float* dWp = a.dparams_partial + (size_t)blockIdx.x * a.part_stride;

  const int nfwd = (a.store_last || a.contract) ? Lp : Lp - 1;

  for (int unit = blockIdx.x; unit < a.tg.n_units; unit += gridDim.x) {
    int n0, n1, kbeg, kend;
    if (NODE) {
      n0 = unit * TE;
      n1 = min(a.tg.N, n0 + TE);
      kbeg = n0;
      kend = n1;
    } else {
      n0 = a.tg.unit_ptr[unit];
      n1 = a.tg.unit_ptr[unit + 1];
      kbeg = a.tg.rowptr[n0];
      kend = a.tg.rowptr[n1];
      if (a.contract == 2) gno_zero_isolated(a.tg.rowptr, n0, n1, (size_t)a.gno_Ka * a.gin, a.gno_S);
      if (a.has_dst_side) {
        for (int item = tid; item < (n1 - n0) * a.dx; item += NT) {
          const int jj = item / a.dx;
          if (a.tg.rowptr[n0 + jj] == a.tg.rowptr[n0 + jj + 1]) a.dxdst[(size_t)n0 * a.dx + item] = 0.f;
        }
      }
    }
    for (int k0 = kbeg; k0 < kend; k0 += TE) {
      const int ne = min(TE, kend - k0);
      load_ids<TE>(a.tg, NODE, k0, ne, s_src, s_dst, s_perm);
      __syncthreads();
      gather_tile<TE>(base + a.zoff[0], a, ne, s_src, s_dst, s_perm);
      __syncthreads();
      // ---- recompute the forward activations of this tile (nothing per-edge was saved by the forward) ----
      for (int l = (a.debug_skip & 16) ? nfwd : 0; l < nfwd; ++l) {
        const float* bias = mlp.b_off[l] >= 0 ? a.params + mlp.b_off[l] : nullptr;
        const float* add = (NODE && l == L - 1) ? a.addend : nullptr;
        dense_tile<TE>(base + a.zoff[l], base + a.zoff[l + 1], mlp.dims[l], mlp.dims[l + 1], a.params + mlp.w_off[l],
                       bias, mlp.act[l], ws, add, k0, ne);
      }
      float* G = base + a.offG0;
      float* Gn = base + a.offG1;

      if (!a.contract) {
        // ---- cotangent of the tile's MLP output ----
        const int d = a.dout;
        const float* ZL = base + a.zoff[Lp];
        for (int item = tid; item < d * TE; item += NT) {
          int e, c;
          if (d >= 16) { e = item / d; c = item - e * d; } else { c = item / TE; e = item - c * TE; }
          float v = 0.f;
          if (e < ne) {
            if (NODE) {
              v = a.gout_ptr[(size_t)(k0 + e) * d + c];
            } else {
              const int dn = s_dst[e];
              v = a.gedge != nullptr ? a.gedge[(size_t)(k0 + e) * d + c] : a.gout_ptr[(size_t)dn * d + c];
              if (a.aggr == NGPDE_AGGR_MEAN) {
                v = __fdiv_rn(v, (float)(a.tg.rowptr[dn + 1] - a.tg.rowptr[dn]));
              } else if (a.aggr == NGPDE_AGGR_MAX || a.aggr == NGPDE_AGGR_MIN) {
                v = (ZL[c * C::LD + e] == a.fwd_out[(size_t)dn * d + c]) ? v : 0.f;
              }
            }
          }
          G[c * C::LD + e] = v;
        }
        __syncthreads();
      } else if (a.contract == 2) {
        // ---- GNOConv, factored (ngpde_gno.cuh): d h_e = T_n' [z_e; 1], d z_e = T_n h_e, S rebuilt for dB = S' DM ----
        const int K = mlp.dims[L - 1], Ka = a.gno_Ka;
        const float* Zin = base + a.zoff[Lp];
        float* Ht = base + a.offH;
        float* Zt = base + a.offZt;
        float* Ts = base + a.offTs;
        const int ldh = a.gin + 4, ldz = gno_ldz(Ka), lds = a.gin + 4;
        const size_t R = (size_t)Ka * a.gin;
        gno_gather_h<TE>(a.arr[ARR_X], a.ld[ARR_X], a.gin, s_src, ne, Ht, ldh);
        gno_transpose_z<TE>(Zin, K, Ka > K, ne, Zt, ldz);
        for (int item = tid; item < K * C::LD; item += NT) G[item] = 0.f;
        // pad rows of the staged T_n (read by the 4-row groups of gno_apply_T); Ts shares the weight staging buffer
        for (int i = tid + Ka * lds; i < ((Ka + 3) & ~3) * lds; i += NT) Ts[i] = 0.f;
        if (a.debug_skip & 32) __syncthreads();
        for (int n = n0; n < n1; ++n) {
          const int r0 = a.tg.rowptr[n], r1 = a.tg.rowptr[n + 1];
          const int lo = max(r0, k0) - k0, hi = min(r1, k0 + ne) - k0;
          if (lo >= hi) continue;
          const float* __restrict__ Tn = a.gno_T + (size_t)n * R;
          if (a.debug_skip & 32) {  // experiment: T_n read straight from global memory (L1), no staging, no barriers
            gno_apply_T<TE, true>(Tn, a.gin, Zt, ldz, Ht, ldh, K, Ka, a.gin, lo, hi, k0, a.desrc, a.dx, G);
            continue;
          }
          __syncthreads();
          const int q = a.gin >> 2;
          if (!(a.debug_skip & 8))
          for (int item = tid; item < Ka * q; item += NT) {
            const int j = item / q, c4 = (item - j * q) * 4;
            *reinterpret_cast<float4*>(Ts + j * lds + c4) = __ldg(reinterpret_cast<const float4*>(Tn + (size_t)j * a.gin + c4));
          }
          __syncthreads();
          if (!(a.debug_skip & 1)) gno_apply_T<TE>(Ts, lds, Zt, ldz, Ht, ldh, K, Ka, a.gin, lo, hi, k0, a.desrc, a.dx, G);
        }
        if (!(a.debug_skip & 2)) gno_outer_rows<TE>(Zt, ldz, Ht, ldh, Ka, a.gin, a.tg.rowptr, n0, n1, k0, ne, a.gno_S);
        __syncthreads();
      } else {
        // ---- GNOConv: backward of m[o] = sum_i act(phi_L(z))[o + gout*i] * h_src[i] ----
        const int l = L - 1;
        const int K = mlp.dims[l], NL = mlp.dims[l + 1];
        const float* W = a.params + mlp.w_off[l];
        const float* bias = mlp.b_off[l] >= 0 ? a.params + mlp.b_off[l] : nullptr;
        const float* Wt = a.wt + mlp.w_off[l];
        const int act = mlp.act[l];
        const float* Zin = base + a.zoff[Lp];
        float* H = base + a.offH;
        float* DM = base + a.offDM;
        float* P = base + a.offP;
        float* DH = base + a.offDH;
        float* red = base + a.offRed;
        float* DZ = G;
        const float* __restrict__ X = a.arr[ARR_X];
        const int ldx = a.ld[ARR_X];
        for (int item = tid; item < a.gin * TE; item += NT) {
          int e, f;
          if (a.gin >= 16) { e = item / a.gin; f = item - e * a.gin; } else { f = item / TE; e = item - f * TE; }
          H[f * C::LD + e] = (e < ne) ? X[(size_t)s_src[e] * ldx + f] : 0.f;
          DH[f * C::LD + e] = 0.f;
        }
        for (int item = tid; item < a.gout * TE; item += NT) {
          int e, c;
          if (a.gout >= 16) { e = item / a.gout; c = item - e * a.gout; } else { c = item / TE; e = item - c * TE; }
          float v = 0.f;
          if (e < ne) {
            const int dn = s_dst[e];
            v = a.gedge != nullptr ? a.gedge[(size_t)(k0 + e) * a.gout + c] : a.gout_ptr[(size_t)dn * a.gout + c];
            if (a.aggr == NGPDE_AGGR_MEAN) v = __fdiv_rn(v, (float)(a.tg.rowptr[dn + 1] - a.tg.rowptr[dn]));
          }
          DM[c * C::LD + e] = v;
        }
        for (int item = tid; item < K * C::LD; item += NT) DZ[item] = 0.f;
        __syncthreads();
        for (int ob = 0; ob * C::NPASS < a.gout; ++ob) {
          const int ncols = min(C::NPASS, a.gout - ob * C::NPASS);
          for (int i = 0; i < a.gin; ++i) {
            const int ncol0 = a.gout * i + ob * C::NPASS;
            auto colfn = [&](int, int& c0, int& nv) {
              c0 = ncol0;
              nv = ncols;
            };
            auto epi = [&](int, float (&acc)[C::RN][C::RE]) {
              float part[C::RE];
#pragma unroll
              for (int q = 0; q < C::RE; ++q) part[q] = 0.f;
#pragma unroll
              for (int j = 0; j < C::RN; ++j) {
                const int ol = n_of<TE>(tn, j);
                if (ol < ncols) {
                  const float b = bias ? bias[ncol0 + ol] : 0.f;
#pragma unroll
                  for (int q = 0; q < C::RE; ++q) {
                    const int e = e_of<TE>(te, q);
                    const float pre = acc[j][q] + b;
                    const float dmv = DM[(ob * C::NPASS + ol) * C::LD + e];
                    part[q] = fmaf(act_fwd(act, pre), dmv, part[q]);
                    float g = dmv * H[i * C::LD + e];
                    if (act != NGPDE_ACT_IDENTITY) g *= act_grad_pre(act, pre);
                    P[ol * C::LD + e] = g;
                  }
                }
              }
#pragma unroll
              for (int q = 0; q < C::RE; ++q) red[tn * TE + e_of<TE>(te, q)] = part[q];
            };
            tile_gemm<TE>(Zin, K, W, NL, 1, colfn, ws, epi);
            if (tid < TE) {
              float s = DH[i * C::LD + tid];
              for (int t = 0; t < C::NTH; ++t) s += red[t * TE + tid];
              DH[i * C::LD + tid] = s;
            }
            tile_outer<TE>(dWp + mlp.w_off[l] + ncol0, NL, mlp.b_off[l] >= 0 ? dWp + mlp.b_off[l] + ncol0 : nullptr,
                           Zin, K, P, ncols);
            if (Lp > 0) {
              const int npass = (K + C::NPASS - 1) / C::NPASS;
              auto colfn2 = [&](int p, int& c0, int& nv) {
                c0 = p * C::NPASS;
                nv = min(C::NPASS, K - c0);
              };
              auto epi2 = [&](int p, float (&acc)[C::RN][C::RE]) {
#pragma unroll
                for (int j = 0; j < C::RN; ++j) {
                  const int k = p * C::NPASS + n_of<TE>(tn, j);
                  if (k < K) {
#pragma unroll
                    for (int q = 0; q < C::RE; ++q) DZ[k * C::LD + e_of<TE>(te, q)] += acc[j][q];
                  }
                }
              };
              tile_gemm<TE>(P, ncols, Wt + (size_t)ncol0 * K, K, npass, colfn2, ws, epi2);
            } else {
              __syncthreads();
            }
          }
        }
        // source-side input gradient of this tile's edges: d h_src
        for (int item = tid; item < a.gin * TE; item += NT) {
          int e, f;
          if (a.gin >= 16) { e = item / a.gin; f = item - e * a.gin; } else { f = item / TE; e = item - f * TE; }
          if (e < ne) a.desrc[(size_t)(k0 + e) * a.dx + f] = DH[f * C::LD + e];
        }
        __syncthreads();
      }

      // ---- back through the Dense layers ----
      for (int l = (a.debug_skip & 4) ? -1 : Lp - 1; l >= 0; --l) {
        const int K = mlp.dims[l], N = mlp.dims[l + 1];
        const float* W = a.params + mlp.w_off[l];
        const float* bias = mlp.b_off[l] >= 0 ? a.params + mlp.b_off[l] : nullptr;
        const int act = mlp.act[l];
        const float* Zl = base + a.zoff[l];
        if (act != NGPDE_ACT_IDENTITY) {
          if (act_grad_from_y(act)) {
            const float* Zo = base + a.zoff[l + 1];
            for (int item = tid; item < N * TE; item += NT) {
              const int n = item / TE, e = item - n * TE;
              G[n * C::LD + e] *= act_grad_y(act, Zo[n * C::LD + e]);
            }
            __syncthreads();
          } else {
            // swish / gelu: the derivative needs the pre-activation -> recompute it and scale G in place
            const float* add = (NODE && l == L - 1) ? a.addend : nullptr;
            const int npass = (N + C::NPASS - 1) / C::NPASS;
            auto colfn = [&](int p, int& c0, int& nv) {
              c0 = p * C::NPASS;
              nv = min(C::NPASS, N - c0);
            };
            auto epi = [&](int p, float (&acc)[C::RN][C::RE]) {
#pragma unroll
              for (int j = 0; j < C::RN; ++j) {
                const int n = p * C::NPASS + n_of<TE>(tn, j);
                if (n < N) {
                  const float b = bias ? bias[n] : 0.f;
#pragma unroll
                  for (int q = 0; q < C::RE; ++q) {
                    const int e = e_of<TE>(te, q);
                    float pre = acc[j][q] + b;
                    if (add != nullptr && e < ne) pre = (acc[j][q] + add[(size_t)(k0 + e) * N + n]) + b;
                    G[n * C::LD + e] *= act_grad_pre(act, pre);
                  }
                }
              }
            };
            tile_gemm<TE>(Zl, K, W, N, npass, colfn, ws, epi);
          }
        }
        if (NODE && a.addend != nullptr && l == L - 1) {
          for (int item = tid; item < ne * N; item += NT) {
            const int e = item / N, n = item - e * N;
            a.dmbar[(size_t)(k0 + e) * N + n] = G[n * C::LD + e];
          }
        }
        tile_outer<TE>(dWp + mlp.w_off[l], N, mlp.b_off[l] >= 0 ? dWp + mlp.b_off[l] : nullptr, Zl, K, G, N);
        if (l > 0 || a.need_dz0) {
          const float* Wt = a.wt + mlp.w_off[l];
          const int npass = (K + C::NPASS - 1) / C::NPASS;
          auto colfn = [&](int p, int& c0, int& nv) {
            c0 = p * C::NPASS;
            nv = min(C::NPASS, K - c0);
          };
          float* Go = Gn;
          auto epi = [&](int p, float (&acc)[C::RN][C::RE]) {
#pragma unroll
            for (int j = 0; j < C::RN; ++j) {
              const int k = p * C::NPASS + n_of<TE>(tn, j);
              if (k < K) {
                float* o = Go + (size_t)k * C::LD + te * 4;
                *reinterpret_cast<float4*>(o) = *reinterpret_cast<float4*>(&acc[j][0]);
                if (C::RE == 8) *reinterpret_cast<float4*>(o + TE / 2) = *reinterpret_cast<float4*>(&acc[j][C::RE - 4]);
              }
            }
          };
          tile_gemm<TE>(G, N, Wt, K, npass, colfn, ws, epi);
          float* t = G; G = Gn; Gn = t;
        }
      }

      // ---- hand the input cotangent dZ0 back to the arrays it was gathered from ----
      if (a.need_dz0) {
        if (NODE) {
          for (int si = 0; si < a.n_segs; ++si) {
            const Seg sg = a.segs[si];
            float* dstp = sg.arr == ARR_X ? a.dx_direct : (sg.arr == ARR_M ? a.dmbar : nullptr);
            if (dstp == nullptr || sg.kind != SEG_DST) continue;
            const int ld = a.ld[sg.arr];
            for (int item = tid; item < ne * sg.width; item += NT) {
              const int e = item / sg.width, f = item - e * sg.width;
              dstp[(size_t)(k0 + e) * ld + sg.col + f] = G[(sg.row + f) * C::LD + e];
            }
          }
        } else {
          const int dx = a.dx;
          // source side: one row per edge, reduced later over the src-sorted transpose
          for (int item = tid; item < dx * TE; item += NT) {
            int e, c;
            if (dx >= 16) { e = item / dx; c = item - e * dx; } else { c = item / TE; e = item - c * TE; }
            if (e >= ne) continue;
            float v = 0.f;
            for (int si = 0; si < a.n_segs; ++si) {
              const Seg sg = a.segs[si];
              if (sg.arr != ARR_X || c < sg.col || c >= sg.col + sg.width) continue;
              const float cf = coef_src(sg.kind);
              if (cf != 0.f) v = fmaf(cf, G[(sg.row + c - sg.col) * C::LD + e], v);
            }
            a.desrc[(size_t)(k0 + e) * dx + c] = v;
          }
          // destination side: sequential over the row's edges, carried across tiles like the forward aggregate
          if (a.has_dst_side) {
            for (int item = tid; item < (n1 - n0) * dx; item += NT) {
              const int jj = item / dx, c = item - jj * dx;
              const int j = n0 + jj;
              const int r0 = a.tg.rowptr[j], r1 = a.tg.rowptr[j + 1];
              const int lo = max(r0, k0), hi = min(r1, k0 + ne);
              if (lo >= hi) continue;
              float accv = (lo == r0) ? 0.f : a.dxdst[(size_t)j * dx + c];
              for (int si = 0; si < a.n_segs; ++si) {
                const Seg sg = a.segs[si];
                if (sg.arr != ARR_X || c < sg.col || c >= sg.col + sg.width) continue;
                const float cf = coef_dst(sg.kind);
                if (cf == 0.f) continue;
                const float* gr = G + (size_t)(sg.row + c - sg.col) * C::LD - k0;
                for (int e = lo; e < hi; ++e) accv = fmaf(cf, gr[e], accv);
              }
              a.dxdst[(size_t)j * dx + c] = accv;
            }
          }
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace ngpde
