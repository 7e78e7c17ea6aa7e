// Tile-level pieces of the factored GNOConv evaluation (ngpde_gno.cuh) used inside the fused edge kernels.
//
// Edge-major staging: Ht[e][i] = x[src(e)][i] (row stride ldh = gin + 4), Zt[e][j] = [z_e; 1; 0...] (row stride ldz,
// ldz % 8 == 4, at least round_up(K + 1, 8) columns) -- both strides keep a float4 row read by 16 lanes with distinct e
// down to two wavefronts.  Thread grid 16 x 8: ti = tid & 15, tj = tid >> 4.
#pragma once
#include "ngpde_tile.cuh"

namespace ngpde {

__host__ __device__ __forceinline__ int gno_ldz(int Ka) {
  const int kp8 = (Ka + 7) & ~7;
  return kp8 + 4;
}

// Zt[e][0:K] = Z[0:K][e] (feature-major tile, row stride LD), Zt[e][K] = 1 when ones != 0, zero padding up to ldz.
template <int TE>
__device__ __forceinline__ void gno_transpose_z(const float* __restrict__ Z, int K, int ones, int ne,
                                                float* __restrict__ Zt, int ldz) {
  using C = Cfg<TE>;
  for (int item = threadIdx.x; item < TE * ldz; item += NT) {
    const int j = item / TE, e = item - j * TE;  // consecutive lanes -> consecutive e: conflict-free read of Z
    float v = 0.f;
    if (e < ne) {
      if (j < K) v = Z[j * C::LD + e];
      else if (j == K && ones) v = 1.f;
    }
    Zt[e * ldz + j] = v;
  }
}

// Ht[e][0:gin] = x[src(e)][0:gin] (zero rows for e >= ne); gin % 4 == 0, ldx % 4 == 0.
template <int TE>
__device__ __forceinline__ void gno_gather_h(const float* __restrict__ X, int ldx, int gin, const int* s_src, int ne,
                                             float* __restrict__ Ht, int ldh) {
  const int q = gin >> 2;
  for (int item = threadIdx.x; item < TE * q; item += NT) {
    const int e = item / q, c4 = (item - e * q) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e < ne) v = __ldg(reinterpret_cast<const float4*>(X + (size_t)s_src[e] * ldx + c4));
    *reinterpret_cast<float4*>(Ht + e * ldh + c4) = v;
  }
}

// S[n][j][i] (+)= sum over the edges of row n inside this tile of Zt[e][j] * Ht[e][i], j < Ka, i < gin, for every
// destination row n0 <= n < n1.  Edges are added in ascending CSR (= original) order; a row continued from an earlier
// tile of the same CTA is carried through S itself (the same thread owns the same entries).  No synchronisation inside.
template <int TE>
__device__ __forceinline__ void gno_outer_rows(const float* __restrict__ Zt, int ldz, const float* __restrict__ Ht,
                                               int ldh, int Ka, int gin, const int* __restrict__ rowptr, int n0, int n1,
                                               int k0, int ne, float* __restrict__ S) {
  const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
  const size_t R = (size_t)Ka * gin;
  for (int n = n0; n < n1; ++n) {
    const int r0 = rowptr[n], r1 = rowptr[n + 1];
    const int lo = max(r0, k0) - k0, hi = min(r1, k0 + ne) - k0;
    if (lo >= hi) continue;
    const bool first = r0 >= k0;
    float* Sn = S + (size_t)n * R;
    for (int jb = 0; jb < Ka; jb += 64) {
      const int j0 = jb + tj * 8;
      if (j0 >= Ka) continue;
      for (int ib = 0; ib < gin; ib += 64) {
        const int i0 = ib + ti * 4;
        if (i0 >= gin) continue;
        float acc[8][4];
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[q][c] = 0.f;
        const float* zp = Zt + lo * ldz + j0;
        const float* hp = Ht + lo * ldh + i0;
#pragma unroll 2
        for (int e = lo; e < hi; ++e, zp += ldz, hp += ldh) {
          const float4 h = *reinterpret_cast<const float4*>(hp);
          float z[8];
          *reinterpret_cast<float4*>(&z[0]) = *reinterpret_cast<const float4*>(zp);
          *reinterpret_cast<float4*>(&z[4]) = *reinterpret_cast<const float4*>(zp + 4);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            acc[q][0] = fmaf(z[q], h.x, acc[q][0]);
            acc[q][1] = fmaf(z[q], h.y, acc[q][1]);
            acc[q][2] = fmaf(z[q], h.z, acc[q][2]);
            acc[q][3] = fmaf(z[q], h.w, acc[q][3]);
          }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (j0 + q < Ka) {
            float4* p = reinterpret_cast<float4*>(Sn + (size_t)(j0 + q) * gin + i0);
            float4 v = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
            if (!first) {
              const float4 o = *p;
              v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *p = v;
          }
        }
      }
    }
  }
}

// Rows of S that belong to isolated destinations are zero.
__device__ __forceinline__ void gno_zero_isolated(const int* __restrict__ rowptr, int n0, int n1, size_t R,
                                                  float* __restrict__ S) {
  for (int n = n0; n < n1; ++n) {
    if (rowptr[n] != rowptr[n + 1]) continue;
    float4* p = reinterpret_cast<float4*>(S + (size_t)n * R);
    for (size_t i = threadIdx.x; i < R / 4; i += NT) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Pullback through the contraction for the edges of ONE destination row n (tile-local range [lo, hi)), given
// T_n = B DM_n staged in shared memory as Ts[j][i] (row stride lds = gin + 4, rows j < round_up(Ka, 4); pad rows zero):
//     desrc[k0 + e][i] = sum_{j < Ka} Zt[e][j] * Ts[j][i]                 (cotangent of h_e = x[src(e)])
//     G[j][e]          = sum_{i < gin} Ts[j][i] * Ht[e][i],   j < K        (cotangent of z_e, feature-major tile)
// 16 edges per pass (lane te = tid & 15), 8 column groups of 8 (tc = tid >> 4); gin % 8 == 0.
// TG: Ts is T_n itself in global memory (row stride gin, no pad rows: row reads are clamped to Ka-1, where za is 0).
template <int TE, bool TG = false>
__device__ __forceinline__ void gno_apply_T(const float* __restrict__ Ts, int lds, const float* __restrict__ Zt, int ldz,
                                            const float* __restrict__ Ht, int ldh, int K, int Ka, int gin, int lo, int hi,
                                            int k0, float* __restrict__ desrc, int dx, float* __restrict__ G) {
  using C = Cfg<TE>;
  const int te = threadIdx.x & 15, tc = threadIdx.x >> 4;
  const int Ka4 = (Ka + 3) & ~3;
  const int jmax = TG ? Ka - 1 : Ka4 - 1;
  auto ld4 = [](const float* p) {
    return TG ? __ldg(reinterpret_cast<const float4*>(p)) : *reinterpret_cast<const float4*>(p);
  };
  for (int eg = lo; eg < hi; eg += 16) {
    const int e = eg + te;
    const bool valid = e < hi;
    const int er = min(e, TE - 1);
    const float* zrow = Zt + er * ldz;
    const float* hrow = Ht + er * ldh;
    // ---- d h ----
    for (int ib = 0; ib < gin; ib += 64) {
      const int i0 = ib + tc * 8;
      if (i0 >= gin) continue;
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
      const float* tp = Ts + i0;
#pragma unroll 4
      for (int j4 = 0; j4 < Ka4; j4 += 4) {
        float z[4];
        *reinterpret_cast<float4*>(&z[0]) = *reinterpret_cast<const float4*>(zrow + j4);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int jr = TG ? min(j4 + jj, jmax) : j4 + jj;
          const float4 t0 = ld4(tp + jr * lds);
          const float4 t1 = ld4(tp + jr * lds + 4);
          acc[0] = fmaf(z[jj], t0.x, acc[0]); acc[1] = fmaf(z[jj], t0.y, acc[1]);
          acc[2] = fmaf(z[jj], t0.z, acc[2]); acc[3] = fmaf(z[jj], t0.w, acc[3]);
          acc[4] = fmaf(z[jj], t1.x, acc[4]); acc[5] = fmaf(z[jj], t1.y, acc[5]);
          acc[6] = fmaf(z[jj], t1.z, acc[6]); acc[7] = fmaf(z[jj], t1.w, acc[7]);
        }
      }
      if (valid) {
        float* o = desrc + (size_t)(k0 + e) * dx + i0;
        *reinterpret_cast<float4*>(o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      }
    }
    // ---- d z ----
    for (int jb = 0; jb < K; jb += 64) {
      const int j0 = jb + tc * 8;
      if (j0 >= K) continue;
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
      const float* tp = Ts + j0 * lds;  // rows j0..j0+7 exist up to Ka4 >= K; rows >= K are never written out
#pragma unroll 4
      for (int i4 = 0; i4 < gin; i4 += 4) {
        const float4 h = *reinterpret_cast<const float4*>(hrow + i4);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 t = ld4(tp + min(q, jmax - j0) * lds + i4);
          acc[q] = fmaf(h.x, t.x, acc[q]);
          acc[q] = fmaf(h.y, t.y, acc[q]);
          acc[q] = fmaf(h.z, t.z, acc[q]);
          acc[q] = fmaf(h.w, t.w, acc[q]);
        }
      }
      if (valid) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          if (j0 + q < K) G[(j0 + q) * C::LD + e] = acc[q];
      }
    }
  }
}

}  // namespace ngpde
