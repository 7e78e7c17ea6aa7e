// Layer-by-layer evaluation of a message-passing layer whose MLPs are too wide for the fused tcgen05 kernels (outputs > 64
// columns: C2's MPPDEConv, hidden 128; ngpde_tc_layout.cuh) -- included by ngpde_conv.cu.
//
// The fused FFMA engine keeps a 64-edge tile on chip through the whole MLP; at 128-wide layers its per-tile chain is latency
// bound (C2: 2.6 ms per RHS, 6 % of the FFMA pipe).  Here every Dense layer is ONE tcgen05 GEMM over all rows (edges in CSR order,
// or nodes) on the 3xTF32 kernel of ngpde_gno_tc.cu, with thin elementwise kernels between them:
//
//   forward   Z0 = assemble(segments)            [rows][ld0]  (ld0 = phi's input width rounded up to 4, pad columns zero)
//             U_l = Z_l Wp_l;  Z_{l+1} = act(U_l + b_l)         (Wp_l = W_l with zero rows up to ld_l: a 16-byte aligned copy)
//             edge phase: mbar[n] = sum / mean over the CSR row of n, ascending (= stored edge order);  node phase: y = Z_L
//   backward  forward once more keeping U_l and Z_l;  G_L = cotangent (edge phase: dmbar[dst] (/ deg));  for l = L-1 .. 0:
//             Gp = G_{l+1} * act'(U_l)  (+ column sums in fixed 64-row chunks -> db_l),  dW_l = Z_l' Gp (split-K GEMM, slices
//             added in order),  G_l = Gp Wp_l';  then dZ0 -> dx (destination side over the CSR row, source side over the
//             transposed row, both ascending) or -> (dx_direct, dmbar) for the node phase.
//
// No atomics anywhere; every reduction has a fixed order.  Activations stay in global memory between the GEMMs (C2: 17 MB per
// layer -- L2-resident on B200); this path trades the fused kernels' on-chip residency for tensor-core rate, which pays
// once layers are >= 96 wide.
#pragma once
#include "ngpde_conv.cuh"
#include "ngpde_gno.cuh"

namespace ngpde {
namespace layered {

inline int r4(int v) { return (v + 3) & ~3; }
inline size_t a256(size_t v) { return (v + 255) & ~size_t(255); }

struct Phase {
  MlpDev mlp;
  int ld[NGPDE_MAX_LAYERS + 1];  // row strides: ld[0] = r4(dims[0]), ld[l] = dims[l] (multiples of 4)
  int woff[NGPDE_MAX_LAYERS];    // float offset of Wp_l ([ld[l]][dims[l+1]]) in the packed image
  int w_floats = 0, maxld = 0, maxwn = 0;
};

inline bool eligible(const MlpDev& m) {
  if (m.L < 1) return false;
  for (int l = 1; l <= m.L; ++l)
    if (m.dims[l] & 3) return false;
  return true;
}

inline Phase make_phase(const MlpDev& m) {
  Phase p{};
  p.mlp = m;
  int off = 0;
  for (int l = 0; l <= m.L; ++l) {
    p.ld[l] = l == 0 ? r4(m.dims[0]) : m.dims[l];
    p.maxld = std::max(p.maxld, p.ld[l]);
    if (l < m.L) {
      p.woff[l] = off;
      off += p.ld[l] * m.dims[l + 1];
      p.maxwn = std::max(p.maxwn, p.ld[l] * m.dims[l + 1]);
    }
  }
  p.w_floats = off;
  return p;
}

constexpr int CHUNK_ROWS = 64;  // rows per block of the fused act' / bias-gradient pass (fixed: the summation order) ...
// ... up to 2048 chunks; beyond, a chunk grows in multiples of 64 rows so the second reduction stage stays short
inline int chunk_rows(int64_t rows) {
  const int64_t per = (rows + 2047) / 2048;
  return (int)std::max<int64_t>(CHUNK_ROWS, (per + CHUNK_ROWS - 1) / CHUNK_ROWS * CHUNK_ROWS);
}

inline int wgrad_splits(const Phase& p, int l, int64_t rows, int num_sms) {
  const int tiles = ((p.ld[l] + 127) / 128) * ((p.mlp.dims[l + 1] + 63) / 64);
  int best = 1;
  double best_cost = 1e30;
  for (int sp = 1; sp <= 1024; ++sp) {  // a 64 x 64 gradient over millions of rows wants hundreds of slices to fill the GPU
    if ((int64_t)sp * 512 > rows && sp > 1) break;
    const int waves = (tiles * sp + 2 * num_sms - 1) / (2 * num_sms);
    const double cost = (double)waves / sp;
    if (cost < best_cost - 1e-12) { best_cost = cost; best = sp; }
  }
  return best;
}

// Byte offsets of one phase's buffers.  KEPT buffers (packed weights, Z_l, U_l: what the backward needs from the forward) are
// relative to `kbase` -- the caller's ngpde_conv_io.state when given, so the backward does not recompute the forward -- the
// scratch buffers relative to the workspace `sbase`.
struct Ws {
  size_t wp = 0, z[NGPDE_MAX_LAYERS + 1] = {}, u[NGPDE_MAX_LAYERS] = {}, kept_end = 0;  // kept
  size_t za = 0, zb = 0, ga = 0, gb = 0, part = 0, cpart = 0, end = 0;                   // scratch
};

// kept part: offsets from `off` in the kept region
inline void plan_kept(const Phase& p, int64_t rows, size_t off, Ws* w) {
  w->wp = off; off = a256(off + 4 * (size_t)p.w_floats);
  for (int l = 0; l < p.mlp.L; ++l) {
    w->z[l] = off; off = a256(off + 4 * (size_t)rows * p.ld[l]);
    w->u[l] = off; off = a256(off + 4 * (size_t)rows * p.ld[l + 1]);
  }
  w->kept_end = off;
}

// scratch part.  Forward without kept buffers: packed weights + two ping-pong activation buffers; forward with kept buffers:
// one buffer for the activated last layer; backward: two cotangent buffers, split-K slices, bias-gradient chunk sums.
inline void plan_scratch(const Phase& p, int64_t rows, bool backward, bool kept, int num_sms, size_t off, Ws* w) {
  if (!backward) {
    if (!kept) { w->wp = off; off = a256(off + 4 * (size_t)p.w_floats); }
    w->za = off; off = a256(off + 4 * (size_t)rows * p.maxld);
    if (!kept) { w->zb = off; off = a256(off + 4 * (size_t)rows * p.maxld); }
  } else {
    w->ga = off; off = a256(off + 4 * (size_t)rows * p.maxld);
    w->gb = off; off = a256(off + 4 * (size_t)rows * p.maxld);
    int maxsp = 1;
    for (int l = 0; l < p.mlp.L; ++l) maxsp = std::max(maxsp, wgrad_splits(p, l, rows, num_sms));
    w->part = off; off = a256(off + 4 * (size_t)maxsp * p.maxwn);
    w->cpart = off; off = a256(off + 4 * (size_t)((rows + chunk_rows(rows) - 1) / chunk_rows(rows)) * p.maxld);
  }
  w->end = off;
}

// ---- kernels ----

// Wp_l[k][n] = W_l[k][n] for k < dims[l], 0 for the pad rows
__global__ void pack_weights_kernel(const float* __restrict__ params, MlpDev m, Phase ph, float* __restrict__ wp) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ph.w_floats; i += gridDim.x * blockDim.x) {
    int l = 0;
    while (l + 1 < m.L && i >= ph.woff[l + 1]) ++l;
    const int j = i - ph.woff[l];
    const int k = j / m.dims[l + 1];
    wp[i] = k < m.dims[l] ? params[m.w_off[l] + j] : 0.f;
  }
}

struct Gather {
  const float* arr[ARR_COUNT];
  int ld[ARR_COUNT];
  int n_segs;
  Seg segs[8];
  const int* src;   // null: node phase (row id = node id)
  const int* dst;
  const int* perm;
  int gdiv;
};

__device__ __forceinline__ float assemble_one(const Gather& g, int f, long long s, long long d, long long pe) {
  for (int si = 0; si < g.n_segs; ++si) {
    const Seg sg = g.segs[si];
    if (f < sg.row || f >= sg.row + sg.width) continue;
    const float* __restrict__ A = g.arr[sg.arr] + sg.col + (f - sg.row);
    const int ld = g.ld[sg.arr];
    switch (sg.kind) {
      case SEG_DST: return A[(size_t)d * ld];
      case SEG_SRC: return A[(size_t)s * ld];
      case SEG_SMD: return A[(size_t)s * ld] - A[(size_t)d * ld];
      case SEG_DMS: return A[(size_t)d * ld] - A[(size_t)s * ld];
      case SEG_EDGE: return A[(size_t)pe * ld];
      default: return A[(size_t)(pe / g.gdiv) * ld];  // SEG_GRAPH
    }
  }
  return 0.f;
}

// Z0[row][f]: the segment list evaluated for one row (edge k in CSR order, or node), pad columns zero.  One thread per four
// columns: a quad inside one segment whose source is 16-byte aligned (`vec` bit of its array) moves as float4.
__global__ void assemble_kernel(Gather g, long long rows, int d0, int ld0, int vec, float* __restrict__ Z) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int q4 = ld0 >> 2;
  if (idx >= rows * q4) return;
  const long long r = idx / q4;
  const int f = (int)(idx - r * q4) * 4;
  const long long s = g.src ? g.src[r] : r, d = g.dst ? g.dst[r] : r, pe = g.perm ? g.perm[r] : r;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  bool done = false;
  for (int si = 0; si < g.n_segs; ++si) {
    const Seg sg = g.segs[si];
    if (f < sg.row || f + 3 >= sg.row + sg.width) continue;
    const int c = sg.col + (f - sg.row);
    if (!((vec >> sg.arr) & 1) || (c & 3)) break;
    const float* __restrict__ A = g.arr[sg.arr] + c;
    const int ld = g.ld[sg.arr];
    auto ld4 = [&](long long row) { return __ldg(reinterpret_cast<const float4*>(A + (size_t)row * ld)); };
    switch (sg.kind) {
      case SEG_DST: v = ld4(d); break;
      case SEG_SRC: v = ld4(s); break;
      case SEG_SMD: { const float4 a = ld4(s), b = ld4(d); v = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); break; }
      case SEG_DMS: { const float4 a = ld4(d), b = ld4(s); v = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); break; }
      case SEG_EDGE: v = ld4(pe); break;
      default: v = ld4(pe / g.gdiv); break;
    }
    done = true;
    break;
  }
  if (!done) {
    if (f < d0) v.x = assemble_one(g, f, s, d, pe);
    if (f + 1 < d0) v.y = assemble_one(g, f + 1, s, d, pe);
    if (f + 2 < d0) v.z = assemble_one(g, f + 2, s, d, pe);
    if (f + 3 < d0) v.w = assemble_one(g, f + 3, s, d, pe);
  }
  *reinterpret_cast<float4*>(Z + (size_t)r * ld0 + f) = v;
}

// u = p (+ add) + b (kept when `U` is given), z = act(u);  P, add, U, Zout are [rows][n]; U / Zout may alias P.  `add` is
// GNOConv's aggregated message entering the node update before the activation (layers.jl:536-547): (W x + mbar) + b.
__global__ void bias_act_kernel(const float* P, const float* __restrict__ bias, int act, long long total, int n,
                                float* U, float* Zout, const float* __restrict__ add = nullptr) {
  const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= total) return;
  const int c = (int)(i4 % n);
  float4 v = *reinterpret_cast<const float4*>(P + i4);
  if (add) {
    const float4 a = *reinterpret_cast<const float4*>(add + i4);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  if (bias) {
    v.x += bias[c]; v.y += bias[c + 1]; v.z += bias[c + 2]; v.w += bias[c + 3];
  }
  if (U) *reinterpret_cast<float4*>(U + i4) = v;
  if (act != NGPDE_ACT_IDENTITY) {
    v.x = act_fwd(act, v.x); v.y = act_fwd(act, v.y); v.z = act_fwd(act, v.z); v.w = act_fwd(act, v.w);
  }
  if (Zout) *reinterpret_cast<float4*>(Zout + i4) = v;
}

__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}

// mbar[n][c] = sum (mean) of M[k][c] over the CSR row of n, ascending k; 0 for an isolated node (NNlib.scatter, layers.jl:111).
// One thread per (node, column quad); d is a multiple of 4.
__global__ void aggregate_rows_kernel(int N, int d, int mean, const int* __restrict__ rowptr, const float* __restrict__ M,
                                      float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int d4 = d >> 2;
  if (idx >= (long long)N * d4) return;
  const int n = (int)(idx / d4), q = (int)(idx - (long long)n * d4);
  const int r0 = rowptr[n], r1 = rowptr[n + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int k = r0; k < r1; ++k) acc = add4(acc, *(reinterpret_cast<const float4*>(M + (size_t)k * d) + q));
  if (mean && r1 > r0) {
    const float deg = (float)(r1 - r0);
    acc = make_float4(__fdiv_rn(acc.x, deg), __fdiv_rn(acc.y, deg), __fdiv_rn(acc.z, deg), __fdiv_rn(acc.w, deg));
  }
  *(reinterpret_cast<float4*>(out + (size_t)n * d) + q) = acc;
}

// Gp[r][c] = Gin[row(r)][c] (/ deg) * act'(U[r][c]);  block b owns rows [crows b, crows (b + 1)) (crows = chunk_rows(rows)): its column sums go to cpart[b][c].
// row_of != null: the edge phase's last layer -- Gin = dmbar is indexed by the edge's destination, divided by the in-degree
// for the mean (true division, like the fused kernels).  256 threads = RG row groups x (n / 4) column quads (float4 accesses,
// n / 4 <= 256; wider rows loop over the quads with RG = 1); a group takes every RG-th row of the chunk in ascending order, the
// groups' sums are added in ascending group order: a fixed summation order for a given n.
__global__ void __launch_bounds__(256) actgrad_kernel(const float* __restrict__ Gin, const float* __restrict__ U, int act,
                                                      long long rows, int n, const int* __restrict__ row_of,
                                                      const int* __restrict__ rowptr_mean, float* __restrict__ Gp,
                                                      float* __restrict__ cpart, int crows) {
  __shared__ float4 sm[256];
  const long long r0 = (long long)blockIdx.x * crows;
  const int nr = (int)((r0 + crows < rows ? r0 + crows : rows) - r0);
  const int n4 = n >> 2;
  const int Q = n4 < 256 ? n4 : 256, RG = 256 / Q;
  const int q0 = threadIdx.x % Q, rg = threadIdx.x / Q;
  for (int qb = 0; qb < n4; qb += Q) {
    const int q = qb + q0;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rg < RG && q < n4) {
#pragma unroll 4
      for (int i = rg; i < nr; i += RG) {
        const long long r = r0 + i;
        float4 g;
        if (row_of) {
          const int dn = row_of[r];
          g = __ldg(reinterpret_cast<const float4*>(Gin + (size_t)dn * n) + q);
          if (rowptr_mean) {
            const float deg = (float)(rowptr_mean[dn + 1] - rowptr_mean[dn]);
            g.x = __fdiv_rn(g.x, deg); g.y = __fdiv_rn(g.y, deg); g.z = __fdiv_rn(g.z, deg); g.w = __fdiv_rn(g.w, deg);
          }
        } else {
          g = *(reinterpret_cast<const float4*>(Gin + (size_t)r * n) + q);
        }
        if (act != NGPDE_ACT_IDENTITY) {
          const float4 u = *(reinterpret_cast<const float4*>(U + (size_t)r * n) + q);
          g.x *= act_grad_pre(act, u.x); g.y *= act_grad_pre(act, u.y); g.z *= act_grad_pre(act, u.z); g.w *= act_grad_pre(act, u.w);
        }
        *(reinterpret_cast<float4*>(Gp + (size_t)r * n) + q) = g;
        acc.x = __fadd_rn(acc.x, g.x); acc.y = __fadd_rn(acc.y, g.y); acc.z = __fadd_rn(acc.z, g.z); acc.w = __fadd_rn(acc.w, g.w);
      }
    }
    if (cpart) {
      sm[threadIdx.x] = acc;
      __syncthreads();
      if (rg == 0 && q < n4) {
        float4 t = sm[q0];
        for (int j = 1; j < RG; ++j) {
          const float4 o = sm[j * Q + q0];
          t.x = __fadd_rn(t.x, o.x); t.y = __fadd_rn(t.y, o.y); t.z = __fadd_rn(t.z, o.z); t.w = __fadd_rn(t.w, o.w);
        }
        *(reinterpret_cast<float4*>(cpart + (size_t)blockIdx.x * n) + q) = t;
      }
      __syncthreads();
    }
  }
}

// db[c] = sum over the chunk sums part[k][c], k < slices: 32 columns x 8 slice lanes per block; lane j adds slices j, j + 8, ...
// in ascending order (loads unrolled), the 8 lane sums are then added in ascending lane order
__global__ void __launch_bounds__(256) colsum_reduce_kernel(const float* __restrict__ part, int slices, int n, float* __restrict__ out) {
  __shared__ float sm[8][33];
  const int cl = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < n) {
#pragma unroll 8
    for (int k = j; k < slices; k += 8) s = __fadd_rn(s, part[(size_t)k * n + c]);
  }
  sm[j][cl] = s;
  __syncthreads();
  if (j == 0 && c < n) {
    float t = sm[0][cl];
    for (int i = 1; i < 8; ++i) t = __fadd_rn(t, sm[i][cl]);
    out[c] = t;
  }
}

// out[i] = sum over the slices s (ascending) of part[s * stride + i], i < P
__global__ void reduce_slices_kernel(const float* __restrict__ part, int slices, size_t stride, int P, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  float s = 0.f;
  for (int k = 0; k < slices; ++k) s = __fadd_rn(s, part[(size_t)k * stride + i]);
  out[i] = s;
}

// edge phase: dx[n][c] = dx_direct[n][c] + sum over the in-edges k of n (ascending) of the destination-side uses of x column c
// in dZ0[k] + sum over the out-edges (transposed row, ascending) of the source-side uses.  One thread per (node, column);
// `edge_dx4_kernel` is the same per column quad for segments that start on multiples of 4 (both in x and in Z0).
__global__ void edge_dx_kernel(Gather g, int N, int dx, int ld0, const int* __restrict__ rowptr, const int* __restrict__ tptr,
                               const int* __restrict__ tpos, const float* __restrict__ dZ0, const float* __restrict__ dx_direct,
                               float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * dx) return;
  const int n = (int)(idx / dx), c = (int)(idx - (long long)n * dx);
  float acc = dx_direct ? dx_direct[idx] : 0.f;
  for (int si = 0; si < g.n_segs; ++si) {
    const Seg sg = g.segs[si];
    if (sg.arr != ARR_X || c < sg.col || c >= sg.col + sg.width) continue;
    const int f = sg.row + (c - sg.col);
    const float cd = coef_dst(sg.kind), cs = coef_src(sg.kind);
    if (cd != 0.f) {
      float s = 0.f;
      for (int k = rowptr[n]; k < rowptr[n + 1]; ++k) s = __fadd_rn(s, dZ0[(size_t)k * ld0 + f]);
      acc = __fadd_rn(acc, cd * s);
    }
    if (cs != 0.f) {
      float s = 0.f;
      for (int q = tptr[n]; q < tptr[n + 1]; ++q) s = __fadd_rn(s, dZ0[(size_t)tpos[q] * ld0 + f]);
      acc = __fadd_rn(acc, cs * s);
    }
  }
  out[idx] = acc;
}

__global__ void edge_dx4_kernel(Gather g, int N, int dx, int ld0, const int* __restrict__ rowptr, const int* __restrict__ tptr,
                                const int* __restrict__ tpos, const float* __restrict__ dZ0, const float* __restrict__ dx_direct,
                                float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int d4 = dx >> 2;
  if (idx >= (long long)N * d4) return;
  const int n = (int)(idx / d4), c = (int)(idx - (long long)n * d4) * 4;
  float4 acc = dx_direct ? *reinterpret_cast<const float4*>(dx_direct + (size_t)n * dx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int si = 0; si < g.n_segs; ++si) {
    const Seg sg = g.segs[si];
    if (sg.arr != ARR_X || c < sg.col || c >= sg.col + sg.width) continue;
    const int f = sg.row + (c - sg.col);
    const float cd = coef_dst(sg.kind), cs = coef_src(sg.kind);
    if (cd != 0.f) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int k = rowptr[n]; k < rowptr[n + 1]; ++k) s = add4(s, *reinterpret_cast<const float4*>(dZ0 + (size_t)k * ld0 + f));
      acc = add4(acc, make_float4(cd * s.x, cd * s.y, cd * s.z, cd * s.w));
    }
    if (cs != 0.f) {
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int q = tptr[n]; q < tptr[n + 1]; ++q) s = add4(s, *reinterpret_cast<const float4*>(dZ0 + (size_t)tpos[q] * ld0 + f));
      acc = add4(acc, make_float4(cs * s.x, cs * s.y, cs * s.z, cs * s.w));
    }
  }
  *reinterpret_cast<float4*>(out + (size_t)n * dx + c) = acc;
}

// true when every x segment starts on a multiple of 4 columns of x and of Z0 and is a multiple of 4 wide
inline bool quads_ok(const Gather& g, int dx) {
  if (dx & 3) return false;
  for (int si = 0; si < g.n_segs; ++si) {
    const Seg& sg = g.segs[si];
    if (sg.arr == ARR_X && ((sg.col | sg.row | sg.width) & 3)) return false;
  }
  return true;
}

// node phase: dZ0 [N][ld0] -> dx_direct [N][dx] (the x segments) and dmbar [N][dm] (the aggregated-message segment)
__global__ void node_split_kernel(Gather g, int N, int dx, int dm, int ld0, const float* __restrict__ dZ0,
                                  float* __restrict__ dx_direct, float* __restrict__ dmbar) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = dx + dm;
  if (idx >= (long long)N * w) return;
  const int n = (int)(idx / w), c = (int)(idx - (long long)n * w);
  const int arr = c < dx ? ARR_X : ARR_M, col = c < dx ? c : c - dx;
  float acc = 0.f;
  for (int si = 0; si < g.n_segs; ++si) {
    const Seg sg = g.segs[si];
    if (sg.arr != arr || sg.kind != SEG_DST || col < sg.col || col >= sg.col + sg.width) continue;
    acc = __fadd_rn(acc, dZ0[(size_t)n * ld0 + sg.row + (col - sg.col)]);
  }
  if (c < dx) dx_direct[(size_t)n * dx + col] = acc;
  else dmbar[(size_t)n * dm + col] = acc;
}

// ---- host ----

inline int gemm(const float* A, int lda, bool a_k, const float* B, int ldb, bool b_k, float* C, int ldc, int64_t M, int N, int64_t K,
                int splits, cudaStream_t st) {
  if (tc_get_enabled() && gno_gemm_tc_supported(lda, a_k, ldb, b_k, ldc, N, K))
    return gno_gemm_tc(A, lda, a_k, B, ldb, b_k, C, ldc, M, N, K, splits, nullptr, st);
  return gno_gemm_ffma(A, lda, a_k, B, ldb, b_k, C, ldc, M, N, K, splits, nullptr, st);
}

inline unsigned blocks(long long total, int per) { return (unsigned)((total + per - 1) / per); }

// forward of one phase.  keep == true: U_l and Z_l are kept (w.u / w.z in kbase) for run_backward.  The activated output of
// the last layer is written to `out` ([rows][dims[L]]) when given, else to a scratch buffer when `last` is asked for; *last
// then points at it.  keep == false: ping-pong buffers in sbase only.
inline int run_forward(const Phase& ph, const Ws& w, char* kbase, char* sbase, const Gather& g, int64_t rows, const float* params,
                       bool keep, float* out, const float** last, cudaStream_t st, const float* addend = nullptr) {
  const MlpDev& m = ph.mlp;
  float* wp = reinterpret_cast<float*>((keep ? kbase : sbase) + w.wp);
  pack_weights_kernel<<<std::min(256u, blocks(ph.w_floats, 256)), 256, 0, st>>>(params, m, ph, wp);
  float* z = reinterpret_cast<float*>(keep ? kbase + w.z[0] : sbase + w.za);
  float* other = keep ? nullptr : reinterpret_cast<float*>(sbase + w.zb);
  int vec = 0;
  for (int a = 0; a < ARR_COUNT; ++a)
    if (g.arr[a] != nullptr && (reinterpret_cast<uintptr_t>(g.arr[a]) & 15) == 0 && (g.ld[a] & 3) == 0) vec |= 1 << a;
  assemble_kernel<<<blocks(rows * (ph.ld[0] / 4), 256), 256, 0, st>>>(g, rows, m.dims[0], ph.ld[0], vec, z);
  for (int l = 0; l < m.L; ++l) {
    const int n = m.dims[l + 1];
    float* u = keep ? reinterpret_cast<float*>(kbase + w.u[l]) : other;
    if (int rc = gemm(z, ph.ld[l], false, wp + ph.woff[l], n, true, u, n, rows, n, ph.ld[l], 1, st)) return rc;
    const float* bias = m.b_off[l] >= 0 ? params + m.b_off[l] : nullptr;
    const long long total = (long long)rows * n;
    const bool lastl = l + 1 == m.L;
    const float* add = lastl ? addend : nullptr;  // enters the last layer before its activation
    if (keep) {
      float* zn = nullptr;
      if (!lastl) zn = reinterpret_cast<float*>(kbase + w.z[l + 1]);
      else if (out) zn = out;
      else if (last && m.act[l] != NGPDE_ACT_IDENTITY) zn = reinterpret_cast<float*>(sbase + w.za);
      if (bias || zn || add) bias_act_kernel<<<blocks(total / 4, 256), 256, 0, st>>>(u, bias, m.act[l], total, n, u, zn, add);
      z = zn ? zn : u;
    } else {
      float* zo = (lastl && out) ? out : u;
      if (bias || add || m.act[l] != NGPDE_ACT_IDENTITY || zo != u)
        bias_act_kernel<<<blocks(total / 4, 256), 256, 0, st>>>(u, bias, m.act[l], total, n, nullptr, zo, add);
      other = z;
      z = zo;
    }
  }
  if (last) *last = z;
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

// backward of one phase after run_forward(keep = true) (in this call, or in the forward call that filled io.state).  Gin: the cotangent of the phase's output -- [rows][dims[L]], or, with
// row_of, dmbar [N][dims[L]] indexed through the edge's destination (mean: divided by the in-degree).  Leaves dZ0 in *dz0
// ([rows][ld[0]]) when need_dz0, and the parameter gradients in dparams (the flat layout of `params`).
inline int run_backward(const Phase& ph, const Ws& w, char* kbase, char* sbase, int64_t rows, const float* Gin, const int* row_of,
                        const int* rowptr_mean, bool need_dz0, int num_sms, float* dparams, const float** dz0, cudaStream_t st,
                        const float** gp0 = nullptr) {
  const MlpDev& m = ph.mlp;
  const float* wp = reinterpret_cast<const float*>(kbase + w.wp);
  float* ga = reinterpret_cast<float*>(sbase + w.ga);
  float* gb = reinterpret_cast<float*>(sbase + w.gb);
  float* part = reinterpret_cast<float*>(sbase + w.part);
  float* cpart = reinterpret_cast<float*>(sbase + w.cpart);
  const int crows = chunk_rows(rows);
  const int nchunks = (int)((rows + crows - 1) / crows);
  const float* g = Gin;
  for (int l = m.L - 1; l >= 0; --l) {
    const int n = m.dims[l + 1];
    const float* u = reinterpret_cast<const float*>(kbase + w.u[l]);
    const float* z = reinterpret_cast<const float*>(kbase + w.z[l]);
    const bool first = l == m.L - 1;
    actgrad_kernel<<<nchunks, 256, 0, st>>>(g, u, m.act[l], rows, n, first ? row_of : nullptr, first ? rowptr_mean : nullptr, ga,
                                            m.b_off[l] >= 0 ? cpart : nullptr, crows);
    if (m.b_off[l] >= 0) colsum_reduce_kernel<<<blocks(n, 32), 256, 0, st>>>(cpart, nchunks, n, dparams + m.b_off[l]);
    const int sp = wgrad_splits(ph, l, rows, num_sms);
    if (int rc = gemm(z, ph.ld[l], true, ga, n, true, part, n, ph.ld[l], n, rows, sp, st)) return rc;
    reduce_slices_kernel<<<blocks(m.dims[l] * n, 256), 256, 0, st>>>(part, sp, (size_t)ph.ld[l] * n, m.dims[l] * n,
                                                                    dparams + m.w_off[l]);
    if (l > 0 || need_dz0) {  // Gp (ga) -> G_l (gb); the next layer's act' pass reads gb and writes ga again
      if (int rc = gemm(ga, n, false, wp + ph.woff[l], n, false, gb, ph.ld[l], rows, ph.ld[l], n, 1, st)) return rc;
      g = gb;
    }
  }
  if (dz0) *dz0 = g;
  if (gp0) *gp0 = ga;  // G * act'(U) of layer 0 (a one-layer phase: the cotangent of an addend to its pre-activation)
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

}  // namespace layered
}  // namespace ngpde
