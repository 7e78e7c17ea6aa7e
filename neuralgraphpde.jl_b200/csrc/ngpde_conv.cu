// Host side of the fused message-passing layers: recipe construction per layer family, shared-memory planning,
// workspace layout, launches, and the extern "C" entry points declared in include/ngpde.h.
#include <algorithm>
#include <cstring>
#include <utility>
#include <vector>

#include "ngpde_conv_launch.cuh"
#include "ngpde_gno_tile.cuh"
#include "ngpde_layered.cuh"
#include "ngpde_gno_node.cuh"
#include "ngpde_tc_layout.cuh"
#include "ngpde_gno.cuh"

namespace ngpde {

// ---- small epilogue kernels ----

// wt[w_off + n*K + k] = params[w_off + k*N + n] for every layer
__global__ void transpose_weights_kernel(const float* __restrict__ params, float* __restrict__ wt, MlpDev mlp) {
  for (int l = 0; l < mlp.L; ++l) {
    const int K = mlp.dims[l], N = mlp.dims[l + 1];
    const float* W = params + mlp.w_off[l];
    float* T = wt + mlp.w_off[l];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K * N; i += gridDim.x * blockDim.x) {
      const int n = i / K, k = i - n * K;
      T[i] = W[(size_t)k * N + n];
    }
  }
}

// dparams[p] = sum over CTAs of partial[cta][p], in ascending CTA order (deterministic)
// aggr = * (layers.jl:49,257,348,441): cotangent of message k of destination n = dmbar[n] * product of the OTHER messages of n
// ([DEP] NNlib's pullback of scatter(*, ...): `prod(j -> src[i, j], inds)` over the destination's other edges in ascending
// order).  Up to 16 in-edges the product is that left fold exactly; longer rows use prefix x suffix products (O(deg)).
// One thread per (destination, channel); `msg` = the recomputed messages [E][d] in CSR order, `suf` scratch of the same size.
__global__ void prod_cotangent_kernel(int N, int d, const int* __restrict__ rowptr, const float* __restrict__ msg,
                                      const float* __restrict__ dmbar, float* __restrict__ suf, float* __restrict__ gedge) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * d) return;
  const int n = (int)(idx / d), c = (int)(idx - (long long)n * d);
  const int r0 = rowptr[n], r1 = rowptr[n + 1];
  const float g = dmbar[idx];
  if (r1 - r0 <= 16) {
    for (int k = r0; k < r1; ++k) {
      float acc = 1.f;
      for (int j = r0; j < r1; ++j)
        if (j != k) acc = __fmul_rn(acc, msg[(size_t)j * d + c]);
      gedge[(size_t)k * d + c] = __fmul_rn(g, acc);
    }
    return;
  }
  float run = 1.f;
  for (int k = r1 - 1; k >= r0; --k) {
    suf[(size_t)k * d + c] = run;
    run = __fmul_rn(msg[(size_t)k * d + c], run);
  }
  run = 1.f;
  for (int k = r0; k < r1; ++k) {
    gedge[(size_t)k * d + c] = __fmul_rn(g, __fmul_rn(run, suf[(size_t)k * d + c]));
    run = __fmul_rn(run, msg[(size_t)k * d + c]);
  }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, int ncta, int P, float* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float s = 0.f;
  for (int c = 0; c < ncta; ++c) s += partial[(size_t)c * P + p];
  out[p] = s;
}

// dx[i][c] = dx_direct[i][c] + dxdst[i][c] + sum over out-edges of i (src-sorted, stable) of desrc[edge][c]
__global__ void dx_combine_kernel(const float* __restrict__ dx_direct, const float* __restrict__ dxdst,
                                  const float* __restrict__ desrc, const int* __restrict__ tptr,
                                  const int* __restrict__ tpos, int N, int dx, int src_c0, int src_w, int dst_c0, int dst_w,
                                  float* __restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * dx) return;
  const int i = (int)(idx / dx), c = (int)(idx - (size_t)i * dx);
  float s = 0.f;
  if (dx_direct) s = dx_direct[idx];
  if (dxdst && c >= dst_c0 && c < dst_c0 + dst_w) s += dxdst[idx];  // columns outside have no destination side (never written)
  if (desrc && c >= src_c0 && c < src_c0 + src_w) {  // desrc is [E][src_w]: the x columns with a source-side use
    for (int q = tptr[i]; q < tptr[i + 1]; ++q) s += desrc[(size_t)tpos[q] * src_w + (c - src_c0)];
  }
  out[idx] = s;
}


// ---- first-layer hoisting (see Plan::hoist below) ----
struct HoistMap {
  Seg segs[8];
  int n_segs;
  int dx, ds;   // node input row r: r < dx -> x column r, else snode column r - dx
  int n1;       // width of phi's first hidden layer
  int w_off0, b_off0, w_off1;  // phi layer 0 weight / bias offsets, offset where layer 1 starts (= end of layer 0)
  int n_params;                // of phi
};

__device__ __forceinline__ int hoist_node_row(const HoistMap& h, const Seg& sg, int f) {
  return sg.arr == ARR_X ? sg.col + f : h.dx + sg.col + f;
}

// phi's first Dense acts on [a_t-side columns; a_s-side columns; differences]: every column is a node quantity, so the layer is
//   W1 z_e + b1 = (Wt' h_t + b1) + Ws' h_s,  h = [x; snode],  Wt = sum_seg coef_dst W1[seg rows], Ws = sum_seg coef_src W1[seg rows]
// out_t: [hdin][n1] then b1 [n1];  out_s: [hdin][n1];  out_in: [n1][n1] identity, then phi's parameters from layer 1 on.
__global__ void hoist_fold_kernel(const float* __restrict__ phi, HoistMap h, float* __restrict__ out_t, float* __restrict__ out_s,
                                  float* __restrict__ out_in) {
  const int hdin = h.dx + h.ds, n1 = h.n1;
  const int total_w = hdin * n1, tail = h.n_params - h.w_off1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total_w + n1 + n1 * n1 + tail; i += gridDim.x * blockDim.x) {
    if (i < total_w) {
      const int nr = i / n1, n = i - nr * n1;
      float wt = 0.f, ws = 0.f;
      for (int si = 0; si < h.n_segs; ++si) {
        const Seg sg = h.segs[si];
        const int f = (sg.arr == ARR_X ? nr : nr - h.dx) - sg.col;
        if ((sg.arr == ARR_X) != (nr < h.dx) || f < 0 || f >= sg.width) continue;
        const float w = phi[h.w_off0 + (size_t)(sg.row + f) * n1 + n];
        wt = fmaf(coef_dst(sg.kind), w, wt);
        ws = fmaf(coef_src(sg.kind), w, ws);
      }
      out_t[i] = wt;
      out_s[i] = ws;
    } else if (i < total_w + n1) {
      const int n = i - total_w;
      out_t[total_w + n] = h.b_off0 >= 0 ? phi[h.b_off0 + n] : 0.f;
    } else if (i < total_w + n1 + n1 * n1) {
      const int j = i - total_w - n1;
      out_in[j] = (j / n1 == j % n1) ? 1.f : 0.f;
    } else {
      const int j = i - total_w - n1 - n1 * n1;
      out_in[n1 * n1 + j] = phi[h.w_off1 + j];
    }
  }
}

// the transpose of the fold: gradients of (Wt, b1, Ws) and of the inner MLP -> dphi
__global__ void hoist_unfold_kernel(HoistMap h, const float* __restrict__ d_t, const float* __restrict__ d_s,
                                    const float* __restrict__ d_in, int din0, float* __restrict__ dphi) {
  const int n1 = h.n1, total_w = (h.dx + h.ds) * n1, tail = h.n_params - h.w_off1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < din0 * n1 + n1 + tail; i += gridDim.x * blockDim.x) {
    if (i < din0 * n1) {
      const int ir = i / n1, n = i - ir * n1;
      float v = 0.f;
      for (int si = 0; si < h.n_segs; ++si) {
        const Seg sg = h.segs[si];
        const int f = ir - sg.row;
        if (f < 0 || f >= sg.width) continue;
        const int nr = hoist_node_row(h, sg, f);
        v = coef_dst(sg.kind) * d_t[(size_t)nr * n1 + n] + coef_src(sg.kind) * d_s[(size_t)nr * n1 + n];
      }
      dphi[h.w_off0 + i] = v;
    } else if (i < din0 * n1 + n1) {
      const int n = i - din0 * n1;
      if (h.b_off0 >= 0) dphi[h.b_off0 + n] = d_t[total_w + n];
    } else {
      const int j = i - din0 * n1 - n1;
      dphi[h.w_off1 + j] = d_in[n1 * n1 + j];
    }
  }
}

// Node phase (gamma over [x; mbar]): W1 [x; m] + b1 = (Wx' x + b1) + Wm' m.  out_u: [dx][n1] then b1; out_v: [dm][n1];
// out_in: identity [n1][n1] then the parameters from layer 1 on.  `row_x` / `row_m`: first input row of the x / mbar segment.
struct NodeHoistMap {
  int dx, dm, n1, row_x, row_m, w_off0, b_off0, w_off1, n_params;
};
__global__ void nhoist_fold_kernel(const float* __restrict__ prm, NodeHoistMap h, float* __restrict__ out_u, float* __restrict__ out_v,
                                   float* __restrict__ out_in) {
  const int n1 = h.n1, nu = h.dx * n1, nv = h.dm * n1, tail = h.n_params - h.w_off1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nu + n1 + nv + n1 * n1 + tail; i += gridDim.x * blockDim.x) {
    if (i < nu) out_u[i] = prm[h.w_off0 + h.row_x * n1 + i];
    else if (i < nu + n1) out_u[i] = h.b_off0 >= 0 ? prm[h.b_off0 + i - nu] : 0.f;
    else if (i < nu + n1 + nv) out_v[i - nu - n1] = prm[h.w_off0 + h.row_m * n1 + (i - nu - n1)];
    else if (i < nu + n1 + nv + n1 * n1) { const int j = i - nu - n1 - nv; out_in[j] = (j / n1 == j % n1) ? 1.f : 0.f; }
    else { const int j = i - nu - n1 - nv - n1 * n1; out_in[n1 * n1 + j] = prm[h.w_off1 + j]; }
  }
}
__global__ void nhoist_unfold_kernel(NodeHoistMap h, const float* __restrict__ d_u, const float* __restrict__ d_v,
                                     const float* __restrict__ d_in, float* __restrict__ dprm) {
  const int n1 = h.n1, nu = h.dx * n1, nv = h.dm * n1, tail = h.n_params - h.w_off1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nu + n1 + nv + tail; i += gridDim.x * blockDim.x) {
    if (i < nu) dprm[h.w_off0 + h.row_x * n1 + i] = d_u[i];
    else if (i < nu + n1) { if (h.b_off0 >= 0) dprm[h.b_off0 + i - nu] = d_u[i]; }
    else if (i < nu + n1 + nv) dprm[h.w_off0 + h.row_m * n1 + (i - nu - n1)] = d_v[i - nu - n1];
    else { const int j = i - nu - n1 - nv; dprm[h.w_off1 + j] = d_in[n1 * n1 + j]; }
  }
}

// Weighted column sums over the node axis in two fixed-order stages: out[j][c] = sum_n S[n][j] dq[n][c] for j < ds, and
// out[ds][c] = sum_n dq[n][c]  (the weight-gradient rows of the static node columns and the bias gradient of a hoisted first
// layer: the dense part goes through the GEMM engine, these few rows do not deserve one).  ds <= 7.
__global__ void __launch_bounds__(256) wcolsum_stage1_kernel(const float* __restrict__ dq, int ld, int C, const float* __restrict__ S, int lds,
                                                              int ds, long long N, int rows_per_block, float* __restrict__ partial) {
  // thread = (4 columns, row slice): the block's rows are dealt round-robin to the slices, whose sums are then added in slice
  // order through shared memory (fixed order: deterministic)
  extern __shared__ float4 sm4[];  // [slices][ds + 1][C / 4]
  const int c4n = C >> 2, slices = blockDim.x / c4n;
  const int c4 = threadIdx.x % c4n, slice = threadIdx.x / c4n;
  const long long r0 = (long long)blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
  float4 acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (slice < slices) {
    for (long long r = r0 + slice; r < r1; r += slices) {
      const float4 v = *reinterpret_cast<const float4*>(dq + (size_t)r * ld + 4 * c4);
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        if (j < ds) {
          const float w = S[(size_t)r * lds + j];
          acc[j].x = fmaf(w, v.x, acc[j].x); acc[j].y = fmaf(w, v.y, acc[j].y);
          acc[j].z = fmaf(w, v.z, acc[j].z); acc[j].w = fmaf(w, v.w, acc[j].w);
        }
      }
      acc[7].x += v.x; acc[7].y += v.y; acc[7].z += v.z; acc[7].w += v.w;
    }
    for (int j = 0; j <= ds; ++j) sm4[(slice * (ds + 1) + j) * c4n + c4] = j < ds ? acc[j] : acc[7];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (ds + 1) * c4n; i += blockDim.x) {
    float4 t = sm4[i];
    for (int sl = 1; sl < slices; ++sl) {
      const float4 u = sm4[sl * (ds + 1) * c4n + i];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    reinterpret_cast<float4*>(partial + (size_t)blockIdx.x * (ds + 1) * C)[i] = t;
  }
}
// out2[g][p] = sum over the `group` partials of group g, in order (first level of a two-level fixed-order reduction)
__global__ void reduce_groups_kernel(const float* __restrict__ partial, int n, int P, int group, float* __restrict__ out2) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x, g = blockIdx.y;
  if (p >= P) return;
  const int c0 = g * group, c1 = min(n, c0 + group);
  float s = 0.f;
  for (int c = c0; c < c1; ++c) s += partial[(size_t)c * P + p];
  out2[(size_t)g * P + p] = s;
}
// [hdin][n1] | [hdin][n1] -> [hdin][2 n1]
__global__ void hoist_cat_kernel(const float* __restrict__ ft, const float* __restrict__ fs, int rows, int n1, float* __restrict__ fcat) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 2 * n1) return;
  const int r = i / (2 * n1), c = i - r * 2 * n1;
  fcat[i] = c < n1 ? ft[r * n1 + c] : fs[r * n1 + c - n1];
}
// dense rows [dx][2 n1] (GEMM) + static / bias rows [(ds + 1)][2 n1] (weighted column sums) -> the (Wt, b1 | Ws) gradient layout
__global__ void hoist_repack_kernel(const float* __restrict__ dwx, const float* __restrict__ dsmall, int dx, int ds, int n1,
                                    float* __restrict__ dft, float* __restrict__ dfs) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, hdin = dx + ds;
  if (i < hdin * n1) {
    const int k = i / n1, c = i - k * n1;
    const float* src = k < dx ? dwx + (size_t)k * 2 * n1 : dsmall + (size_t)(k - dx) * 2 * n1;
    dft[i] = src[c];
    dfs[i] = src[n1 + c];
  } else if (i < hdin * n1 + n1) {
    dft[i] = dsmall[(size_t)ds * 2 * n1 + (i - hdin * n1)];
  }
}

// q[n] = [a[n]; b[n]]  (rows of `w` floats, w % 4 == 0) and its inverse
__global__ void hoist_join_kernel(const float4* __restrict__ a, const float4* __restrict__ b, size_t rows, int w4, float4* __restrict__ q) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 2 * w4) return;
  const size_t r = i / (2 * w4);
  const int c = (int)(i - r * 2 * w4);
  q[i] = c < w4 ? a[r * w4 + c] : b[r * w4 + c - w4];
}
__global__ void hoist_split_kernel(const float4* __restrict__ q, size_t rows, int w4, float4* __restrict__ a, float4* __restrict__ b) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 2 * w4) return;
  const size_t r = i / (2 * w4);
  const int c = (int)(i - r * 2 * w4);
  if (c < w4) a[r * w4 + c] = q[i]; else b[r * w4 + c - w4] = q[i];
}
// out = a + b + c (a and c may be NULL)
__global__ void add3_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c, size_t n, float* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (a ? a[i] : 0.f) + b[i] + (c ? c[i] : 0.f);
}

// the same with four columns per thread (dx, src_c0, src_w, dst_c0, dst_w all multiples of 4; 16-byte aligned arrays): the hoisted
// dQ rows are 128 floats wide and the scalar kernel above spent its time issuing loads
__global__ void dx_combine4_kernel(const float4* __restrict__ dx_direct, const float4* __restrict__ dxdst, const float4* __restrict__ desrc,
                                   const int* __restrict__ tptr, const int* __restrict__ tpos, int N, int dx4, int src_c04, int src_w4,
                                   int dst_c04, int dst_w4, float4* __restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * dx4) return;
  const int i = (int)(idx / dx4), c = (int)(idx - (size_t)i * dx4);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (dx_direct) s = dx_direct[idx];
  if (dxdst && c >= dst_c04 && c < dst_c04 + dst_w4) {
    const float4 t = dxdst[idx];
    s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
  }
  if (desrc && c >= src_c04 && c < src_c04 + src_w4) {
    for (int q = tptr[i]; q < tptr[i + 1]; ++q) {
      const float4 t = __ldg(desrc + (size_t)tpos[q] * src_w4 + (c - src_c04));
      s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
  }
  out[idx] = s;
}

namespace {

// ---- optional per-kernel timing (ngpde_profile_*): CUDA events recorded on the launching stream around the four
// fused kernels, so a benchmark can attribute time to the dominant kernel without a profiler attached ----
struct ProfSlot {
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
};
bool g_prof_on = false;
bool g_gno_factored = true;  // NGPDE_OPT_GNO_FACTORED
int g_debug_skip = 0;          // NGPDE_OPT_DEBUG_SKIP
ProfSlot g_prof[NGPDE_PROF_SLOTS];

struct ProfScope {
  cudaStream_t st;
  cudaEvent_t e1 = nullptr;
  bool on;
  ProfScope(int slot, cudaStream_t s) : st(s), on(g_prof_on) {
    if (!on) return;
    cudaEvent_t e0;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { on = false; return; }
    cudaEventRecord(e0, st);
    g_prof[slot].ev.emplace_back(e0, e1);
  }
  ~ProfScope() {
    if (on) cudaEventRecord(e1, st);
  }
};

// host copy of the sign table coef_dst (the device versions live in the kernels header)
int coef_dst_host(int kind) { return kind == SEG_DST || kind == SEG_SMD || kind == SEG_DMS; }

constexpr int kSmemTwoCtas = 113 * 1024;  // <= this: two CTAs per SM
constexpr int kSmemMax = 227 * 1024;

int make_mlp(const ngpde_mlp& m, MlpDev* out, const char* what) {
  NGPDE_REQUIRE(m.n_layers >= 0 && m.n_layers <= NGPDE_MAX_LAYERS, "%s: n_layers=%d out of range", what, m.n_layers);
  out->L = m.n_layers;
  int off = 0;
  for (int l = 0; l <= m.n_layers; ++l) {
    NGPDE_REQUIRE(m.n_layers == 0 || m.dims[l] > 0, "%s: dims[%d]=%d must be positive", what, l, m.dims[l]);
    out->dims[l] = m.dims[l];
  }
  for (int l = 0; l < m.n_layers; ++l) {
    NGPDE_REQUIRE(m.act[l] >= NGPDE_ACT_IDENTITY && m.act[l] <= NGPDE_ACT_LEAKYRELU, "%s: unknown activation %d", what,
                  m.act[l]);
    out->act[l] = m.act[l];
    out->w_off[l] = off;
    off += m.dims[l] * m.dims[l + 1];
    if (m.has_bias[l]) {
      out->b_off[l] = off;
      off += m.dims[l + 1];
    } else {
      out->b_off[l] = -1;
    }
  }
  out->n_params = off;
  return NGPDE_OK;
}

struct Plan {
  Seg esegs[8];
  int n_esegs = 0;
  Seg nsegs[8];
  int n_nsegs = 0;
  MlpDev phi{}, node{};
  int contract = 0;  // 1: per-edge GNO contraction, 2: factored (ngpde_gno.cuh)
  int gno_K = 0, gno_Ka = 0;  // factored: width of phi's last hidden layer, +1 with a bias row
  int dm = 0;  // width of the aggregated message
  int dy = 0;  // width of the layer output
  int ds = 0;
  bool has_node = false;
  bool node_addend = false;
  bool edge_need_dz0 = false;
  bool edge_dst_side = false;
  // First-layer hoisting.  When phi's first Dense is too wide for the tensor-core kernels (C5: 130 columns, limit 80) and
  // every input column is a node quantity, W1 z_e + b1 = Pt[dst] + Ps[src] with two per-NODE projections (1/deg of the
  // work); the edge kernels then run the INNER problem: x' = Q = [Pt | Ps] ([N][2 n1]), input = Q[dst][0:n1] + Q[src][n1:2n1]
  // (two segments over the same input rows, summed by the gather), MLP = [identity n1 x n1, phi's first activation] followed
  // by phi's layers 1.. -- the unchanged tcgen05 kernels, whose cotangent machinery returns dQ.
  bool hoist = false;
  MlpDev phi_in{}, mlp_t{}, mlp_s{};
  Seg hsegs[2];
  int h_n1 = 0, h_din = 0;
  // the same for the node update over [x; mbar] (C5's gamma: 128 columns): U = x Wx + b1, V = mbar Wm, inner input U + V
  bool nhoist = false;
  MlpDev node_in{}, mlp_u{}, mlp_v{};
  int nh_n1 = 0, nh_row_x = 0, nh_row_m = 0;
  // Layer-by-layer evaluation on the tcgen05 GEMM (ngpde_layered.cuh): MLPs with a layer wider than the fused tensor-core
  // kernels take (outputs > 64), + / mean aggregation, enough edges to fill the GEMM grid
  bool layered = false;
  // Factored GNOConv with phi's hidden layers on the tcgen05 GEMMs (ngpde_layered.cuh) and the per-destination products in a
  // warp-per-node kernel (ngpde_gno_node.cuh) instead of the fused FFMA edge kernel
  bool gno_layered = false;
  bool gno_node_gemm = false;  // ... and its node update act(W x + mbar + b) as one GEMM + an elementwise pass
  MlpDev phi_hidden{};
};
bool g_hoist = true;    // NGPDE_OPT_HOIST
bool g_layered = true;  // NGPDE_OPT_LAYERED
bool g_gno_layered = true;  // NGPDE_OPT_GNO_LAYERED

void push_seg(Seg* segs, int* n, int* row, int kind, int arr, int col, int width) {
  if (width <= 0) return;
  segs[*n] = Seg{kind, arr, col, width, *row};
  *row += width;
  ++*n;
}

int one_layer_mlp(int din, int dout, bool bias, MlpDev* m) {
  ngpde_mlp h{};
  h.n_layers = 1;
  h.dims[0] = din;
  h.dims[1] = dout;
  h.act[0] = NGPDE_ACT_IDENTITY;
  h.has_bias[0] = bias ? 1 : 0;
  return make_mlp(h, m, "hoisted projection");
}

// inner MLP of a hoisted first layer: [identity n1 x n1 with the first activation] + layers 1.., parameters = [I | tail]
MlpDev hoist_inner_mlp(const MlpDev& m) {
  MlpDev in{};
  const int n1 = m.dims[1];
  in.L = m.L;
  in.dims[0] = n1;
  for (int l = 1; l <= in.L; ++l) in.dims[l] = m.dims[l];
  in.act[0] = m.act[0];
  in.w_off[0] = 0;
  in.b_off[0] = -1;
  const int shift = n1 * n1 - m.w_off[1];
  for (int l = 1; l < in.L; ++l) {
    in.act[l] = m.act[l];
    in.w_off[l] = m.w_off[l] + shift;
    in.b_off[l] = m.b_off[l] >= 0 ? m.b_off[l] + shift : -1;
  }
  in.n_params = m.n_params + shift;
  return in;
}

// decides Plan::hoist and fills the inner problem
void plan_hoist(const ngpde_conv_desc& d, Plan* p) {
  p->hoist = false;
  if (!g_hoist || p->contract || p->phi.L < 2 || p->phi.L > NGPDE_MAX_LAYERS) return;
  if (!(d.aggr == NGPDE_AGGR_SUM || d.aggr == NGPDE_AGGR_MEAN)) return;
  for (int i = 0; i < p->n_esegs; ++i) {
    const Seg& sg = p->esegs[i];
    if (!(sg.arr == ARR_X || sg.arr == ARR_S)) return;          // edge / graph features are not node quantities
    if (sg.kind == SEG_EDGE || sg.kind == SEG_GRAPH) return;
  }
  const int din0 = p->phi.dims[0], n1 = p->phi.dims[1], hdin = d.dx + p->ds;
  if ((din0 + 15) / 16 * 16 <= 80) return;                      // narrow enough for the tensor-core kernels as it is
  if (n1 > TC_MAXN || (n1 & 3) != 0 || hdin > 80) return;
  TcBwdPhase tb;
  TcLayout lay;
  int a, b, c, e;
  {  // the original MLP must be off the tensor-core path, the inner one on it (forward and backward)
    if (tc_make_layout(p->phi, 0, false, p->dm, false, &lay, &a, &b, &c, &e) && tc_bwd_make(p->phi, 0, false, false, d.aggr, true, &tb)) return;
  }
  const MlpDev in = hoist_inner_mlp(p->phi);
  if (!tc_make_layout(in, 0, false, p->dm, false, &lay, &a, &b, &c, &e)) return;
  if (!tc_bwd_make(in, 0, false, false, d.aggr, true, &tb)) return;
  if (one_layer_mlp(hdin, n1, true, &p->mlp_t) || one_layer_mlp(hdin, n1, false, &p->mlp_s)) return;
  if (!tc_make_layout(p->mlp_t, 0, false, n1, true, &lay, &a, &b, &c, &e) || !tc_bwd_make(p->mlp_t, 0, false, true, d.aggr, true, &tb)) return;
  p->phi_in = in;
  p->h_n1 = n1;
  p->h_din = hdin;
  p->hsegs[0] = Seg{SEG_DST, ARR_X, 0, n1, 0};
  p->hsegs[1] = Seg{SEG_SRC, ARR_X, n1, n1, 0};
  p->hoist = true;
}

void plan_nhoist(const ngpde_conv_desc& d, Plan* p) {
  p->nhoist = false;
  if (!g_hoist || !p->has_node || p->node_addend || p->node.L < 2 || p->node.L > NGPDE_MAX_LAYERS) return;
  if (p->n_nsegs != 2) return;
  int row_x = -1, row_m = -1;
  for (int i = 0; i < 2; ++i) {
    const Seg& sg = p->nsegs[i];
    if (sg.kind != SEG_DST || sg.col != 0) return;
    if (sg.arr == ARR_X && sg.width == d.dx) row_x = sg.row;
    else if (sg.arr == ARR_M && sg.width == p->dm) row_m = sg.row;
    else return;
  }
  if (row_x < 0 || row_m < 0) return;
  const int din0 = p->node.dims[0], n1 = p->node.dims[1];
  if ((din0 + 15) / 16 * 16 <= 80) return;
  if (n1 > TC_MAXN || (n1 & 3) != 0 || d.dx > 80 || p->dm > 80) return;
  TcBwdPhase tb;
  TcLayout lay;
  int a, b, c, e;
  if (tc_make_layout(p->node, 0, false, p->dy, true, &lay, &a, &b, &c, &e) && tc_bwd_make(p->node, 0, false, true, d.aggr, true, &tb)) return;
  const MlpDev in = hoist_inner_mlp(p->node);
  if (!tc_make_layout(in, 0, false, p->dy, true, &lay, &a, &b, &c, &e)) return;
  if (!tc_bwd_make(in, 0, false, true, d.aggr, true, &tb)) return;
  if (one_layer_mlp(d.dx, n1, true, &p->mlp_u) || one_layer_mlp(p->dm, n1, false, &p->mlp_v)) return;
  if (!tc_make_layout(p->mlp_u, 0, false, n1, true, &lay, &a, &b, &c, &e) || !tc_bwd_make(p->mlp_u, 0, false, true, d.aggr, true, &tb)) return;
  if (!tc_make_layout(p->mlp_v, 0, false, n1, true, &lay, &a, &b, &c, &e) || !tc_bwd_make(p->mlp_v, 0, false, true, d.aggr, true, &tb)) return;
  p->node_in = in;
  p->nh_n1 = n1;
  p->nh_row_x = row_x;
  p->nh_row_m = row_m;
  p->nhoist = true;
}

NodeHoistMap nhoist_map(const ngpde_conv_desc& d, const Plan& p) {
  NodeHoistMap h{};
  h.dx = d.dx; h.dm = p.dm; h.n1 = p.nh_n1; h.row_x = p.nh_row_x; h.row_m = p.nh_row_m;
  h.w_off0 = p.node.w_off[0]; h.b_off0 = p.node.b_off[0]; h.w_off1 = p.node.w_off[1]; h.n_params = p.node.n_params;
  return h;
}

HoistMap hoist_map(const ngpde_conv_desc& d, const Plan& p) {
  HoistMap h{};
  h.n_segs = p.n_esegs;
  std::memcpy(h.segs, p.esegs, sizeof(p.esegs));
  h.dx = d.dx;
  h.ds = p.ds;
  h.n1 = p.h_n1;
  h.w_off0 = p.phi.w_off[0];
  h.b_off0 = p.phi.b_off[0];
  h.w_off1 = p.phi.w_off[1];
  h.n_params = p.phi.n_params;
  return h;
}

int make_plan(const ngpde_graph* g, const ngpde_conv_desc& d, Plan* p) {
  NGPDE_REQUIRE(g != nullptr, "null graph handle");
  NGPDE_REQUIRE(d.dx > 0, "dx must be positive");
  NGPDE_REQUIRE(d.dhs >= 0 && d.dpos >= 0 && d.de >= 0 && d.dtheta >= 0, "negative feature width");
  NGPDE_REQUIRE(d.aggr >= NGPDE_AGGR_SUM && d.aggr <= NGPDE_AGGR_PROD, "unknown aggregation %d", d.aggr);
  if (int rc = make_mlp(d.phi, &p->phi, "phi")) return rc;
  if (int rc = make_mlp(d.node, &p->node, "node")) return rc;
  NGPDE_REQUIRE(p->phi.L >= 1, "phi needs at least one Dense layer");
  p->ds = d.dhs + d.dpos;
  int row = 0, nrow = 0;
  switch (d.family) {
    case NGPDE_EXPLICIT_EDGE_CONV:  // layers.jl:104-106: vcat(hi..., hj..., posj - posi)
      push_seg(p->esegs, &p->n_esegs, &row, SEG_DST, ARR_X, 0, d.dx);
      push_seg(p->esegs, &p->n_esegs, &row, SEG_DST, ARR_S, 0, d.dhs);
      push_seg(p->esegs, &p->n_esegs, &row, SEG_SRC, ARR_X, 0, d.dx);
      push_seg(p->esegs, &p->n_esegs, &row, SEG_SRC, ARR_S, 0, d.dhs);
      push_seg(p->esegs, &p->n_esegs, &row, SEG_SMD, ARR_S, d.dhs, d.dpos);
      p->dm = p->phi.dims[p->phi.L];
      p->dy = p->dm;
      p->has_node = false;
      NGPDE_REQUIRE(p->node.L == 0, "ExplicitEdgeConv has no node update");
      break;
    case NGPDE_VMH_CONV:  // layers.jl:314-316: vcat(hi..., (hj .- hi)..., posj .- posi); :328: gamma(vcat(x..., m))
      push_seg(p->esegs, &p->n_esegs, &row, SEG_DST, ARR_X, 0, d.dx);
      push_seg(p->esegs, &p->n_esegs, &row, SEG_DST, ARR_S, 0, d.dhs);
      push_seg(p->esegs, &p->n_esegs, &row, SEG_SMD, ARR_X, 0, d.dx);
      push_seg(p->esegs, &p->n_esegs, &row, SEG_SMD, ARR_S, 0, d.dhs);
      push_seg(p->esegs, &p->n_esegs, &row, SEG_SMD, ARR_S, d.dhs, d.dpos);
      p->dm = p->phi.dims[p->phi.L];
      push_seg(p->nsegs, &p->n_nsegs, &nrow, SEG_DST, ARR_X, 0, d.dx);
      push_seg(p->nsegs, &p->n_nsegs, &nrow, SEG_DST, ARR_M, 0, p->dm);
      p->has_node = true;
      break;
    case NGPDE_MPPDE_CONV:  // layers.jl:409-410: vcat(hi, hj, di .- dj, e, theta); :418: psi(vcat(x, m, theta))
      push_seg(p->esegs, &p->n_esegs, &row, SEG_DST, ARR_X, 0, d.dx);
      push_seg(p->esegs, &p->n_esegs, &row, SEG_SRC, ARR_X, 0, d.dx);
      push_seg(p->esegs, &p->n_esegs, &row, SEG_DMS, ARR_S, 0, p->ds);
      push_seg(p->esegs, &p->n_esegs, &row, SEG_EDGE, ARR_E, 0, d.de);
      push_seg(p->esegs, &p->n_esegs, &row, SEG_GRAPH, ARR_T, 0, d.dtheta);
      p->dm = p->phi.dims[p->phi.L];
      push_seg(p->nsegs, &p->n_nsegs, &nrow, SEG_DST, ARR_X, 0, d.dx);
      push_seg(p->nsegs, &p->n_nsegs, &nrow, SEG_DST, ARR_M, 0, p->dm);
      push_seg(p->nsegs, &p->n_nsegs, &nrow, SEG_GRAPH, ARR_T, 0, d.dtheta);
      p->has_node = true;
      if (d.dtheta > 0) {
        NGPDE_REQUIRE(g->G > 0 && g->E % g->G == 0 && g->N % g->G == 0,
                      "MPPDEConv assumes equal-sized graphs (E=%lld, N=%lld, G=%lld)", (long long)g->E,
                      (long long)g->N, (long long)g->G);
      }
      break;
    case NGPDE_GNO_CONV:  // layers.jl:516-530, 536-547
      push_seg(p->esegs, &p->n_esegs, &row, SEG_DST, ARR_S, 0, p->ds);
      push_seg(p->esegs, &p->n_esegs, &row, SEG_SRC, ARR_S, 0, p->ds);
      push_seg(p->esegs, &p->n_esegs, &row, SEG_EDGE, ARR_E, 0, d.de);
      NGPDE_REQUIRE(d.gno_in == d.dx && d.gno_out > 0, "GNOConv: in_chs=%d must equal dx=%d", d.gno_in, d.dx);
      NGPDE_REQUIRE(p->phi.dims[p->phi.L] == d.gno_in * d.gno_out, "GNOConv: phi must output in*out=%d values, got %d",
                    d.gno_in * d.gno_out, p->phi.dims[p->phi.L]);
      p->contract = 1;
      {
        // factored evaluation: affine last layer, linear aggregation, vectorisable widths (else the per-edge engine)
        const int l = p->phi.L - 1;
        const bool bias_adjacent = p->phi.b_off[l] < 0 || p->phi.b_off[l] == p->phi.w_off[l] + p->phi.dims[l] * p->phi.dims[l + 1];
        if (g_gno_factored && p->phi.act[l] == NGPDE_ACT_IDENTITY && (d.aggr == NGPDE_AGGR_SUM || d.aggr == NGPDE_AGGR_MEAN) &&
            d.gno_in % 8 == 0 && d.gno_out % 4 == 0 && g->E > 0 && bias_adjacent) {
          p->contract = 2;
          p->gno_K = p->phi.dims[l];
          p->gno_Ka = p->gno_K + (p->phi.b_off[l] >= 0 ? 1 : 0);
        }
      }
      p->dm = d.gno_out;
      push_seg(p->nsegs, &p->n_nsegs, &nrow, SEG_DST, ARR_X, 0, d.dx);
      NGPDE_REQUIRE(p->node.L == 1 && p->node.dims[1] == d.gno_out, "GNOConv: node must be the 1-layer linear map");
      p->has_node = true;
      p->node_addend = true;
      break;
    default:
      set_error("unknown layer family %d", d.family);
      return NGPDE_ERR_INVALID;
  }
  NGPDE_REQUIRE(row == p->phi.dims[0], "phi expects %d input rows but the layer assembles %d (DimensionMismatch)",
                p->phi.dims[0], row);
  if (p->has_node) {
    NGPDE_REQUIRE(nrow == p->node.dims[0], "node update expects %d input rows but the layer assembles %d",
                  p->node.dims[0], nrow);
    p->dy = p->node.dims[p->node.L];
  }
  for (int i = 0; i < p->n_esegs; ++i) {
    if (p->esegs[i].arr == ARR_X) {
      p->edge_need_dz0 = true;
      if (coef_dst_host(p->esegs[i].kind)) p->edge_dst_side = true;
    }
  }
  plan_hoist(d, p);
  plan_nhoist(d, p);
  {
    int widest = 0;
    for (int l = 1; l <= p->phi.L; ++l) widest = std::max(widest, p->phi.dims[l]);
    p->layered = g_layered && tc_get_enabled() && !p->contract && !p->hoist && !p->nhoist && widest > 64 && g->E >= 8192 &&
                 (d.aggr == NGPDE_AGGR_SUM || d.aggr == NGPDE_AGGR_MEAN) && layered::eligible(p->phi) &&
                 (!p->has_node || layered::eligible(p->node));
  }
  if (p->contract == 2 && g_gno_layered && p->phi.L >= 2 && gnonode::supported(p->gno_K, d.gno_in) && g->E >= 2048) {
    p->phi_hidden = p->phi;
    p->phi_hidden.L = p->phi.L - 1;
    p->gno_layered = layered::eligible(p->phi_hidden);
    p->gno_node_gemm = p->gno_layered && p->node_addend && p->node.L == 1 && layered::eligible(p->node) && (d.dx & 3) == 0;
  }
  return NGPDE_OK;
}

struct FwdSmem {
  int offA, offB, offW, offH, offZt, floats;
};

FwdSmem fwd_smem(const MlpDev& m, int contract, int gin, int gout, int te, int Ka = 0) {
  const int ld = te + 4;
  const int npass = (te == 32) ? 128 : 64;
  const int Lp = contract ? m.L - 1 : m.L;
  int rows[2] = {0, 0};
  for (int l = 0; l <= Lp; ++l) rows[l & 1] = std::max(rows[l & 1], m.dims[l]);
  if (contract == 1) rows[(Lp + 1) & 1] = std::max(rows[(Lp + 1) & 1], gout);
  FwdSmem s;
  s.offA = 0;
  s.offB = rows[0] * ld;
  s.offH = s.offB + rows[1] * ld;
  s.offZt = s.offH + (contract == 2 ? te * (gin + 4) : (contract ? gin * ld : 0));
  s.offW = s.offZt + (contract == 2 ? te * gno_ldz(Ka) : 0);
  s.floats = 3 * te + s.offW + 2 * KC * npass;
  return s;
}

struct BwdSmem {
  int zoff[NGPDE_MAX_LAYERS + 1];
  int offG0, offG1, offW, offH, offDM, offP, offDH, offRed, offZt, offTs, floats;
  int store_last;
};

BwdSmem bwd_smem(const MlpDev& m, int contract, int gin, int gout, int aggr, bool node, bool need_dz0, int te, int Ka = 0) {
  const int ld = te + 4;
  const int npass = (te == 32) ? 128 : 64;
  const int nth = (te == 32) ? 16 : 8;
  const int Lp = contract ? m.L - 1 : m.L;
  BwdSmem s{};
  const int last_act = m.act[m.L - 1];
  s.store_last = 0;
  if (!contract) {
    if (last_act != NGPDE_ACT_IDENTITY && act_grad_from_y(last_act)) s.store_last = 1;
    if (!node && (aggr == NGPDE_AGGR_MAX || aggr == NGPDE_AGGR_MIN)) s.store_last = 1;
  }
  int off = 0;
  const int nstore = (s.store_last || contract) ? Lp : Lp - 1;
  for (int l = 0; l <= Lp; ++l) {
    s.zoff[l] = off;
    if (l <= nstore) off += m.dims[l] * ld;
  }
  // G ping-pong: G0 holds dZ_Lp, dZ_{Lp-2}, ...; G1 holds dZ_{Lp-1}, ...
  int rows[2] = {0, 0};
  for (int l = Lp; l >= 0; --l) {
    if (l == 0 && !need_dz0 && Lp > 0) break;
    rows[(Lp - l) & 1] = std::max(rows[(Lp - l) & 1], m.dims[l]);
  }
  rows[0] = std::max(rows[0], m.dims[Lp]);
  s.offG0 = off;
  off += rows[0] * ld;
  s.offG1 = off;
  off += rows[1] * ld;
  int ws_floats = 2 * KC * npass;
  if (contract == 2) {
    // the edge-major h tile is dead before the MLP backward first writes G1, and the staged T_n lives only between the
    // recompute and the MLP backward (the two users of the weight staging buffer): both are aliased, which is what lets a
    // 64-edge tile of the C4 shape fit two CTAs per SM
    s.offH = s.offG1;
    off = s.offG1 + std::max(rows[1] * ld, te * (gin + 4));
    s.offZt = off; off += te * gno_ldz(Ka);
    ws_floats = std::max(ws_floats, ((Ka + 3) & ~3) * (gin + 4));
  } else if (contract) {
    s.offH = off;  off += gin * ld;
    s.offDH = off; off += gin * ld;
    s.offDM = off; off += gout * ld;
    s.offP = off;  off += npass * ld;
    s.offRed = off; off += nth * te;
  }
  s.offW = off;
  s.offTs = off;
  off += ws_floats;
  s.floats = 3 * te + off;
  return s;
}

template <class F>
int pick_tile(F bytes_for, int* te_out, int* bytes_out) {
  const char* cap = getenv("NGPDE_MAX_TE");  // developer switch: cap the FFMA engine's tile size (32 / 64 / 128)
  const int tmax = cap ? (atoi(cap) <= 32 ? 0 : (atoi(cap) <= 64 ? 1 : 2)) : 2;
  for (int pass = 0; pass < 2; ++pass) {
    const int limit = pass == 0 ? kSmemTwoCtas : kSmemMax;
    for (int t = tmax; t >= 0; --t) {
      const int b = bytes_for(kTileSizes[t]);
      if (b <= limit) {
        *te_out = kTileSizes[t];
        *bytes_out = b;
        return NGPDE_OK;
      }
    }
  }
  set_error("layer too wide for the shared-memory tile (needs %d bytes at the smallest tile)", bytes_for(32));
  return NGPDE_ERR_UNSUPPORTED;
}

int tile_index(int te) { return te == 32 ? 0 : (te == 64 ? 1 : 2); }

void fill_arrays(const ngpde_graph* g, const ngpde_conv_desc& d, const Plan& p, const ngpde_conv_io& io,
                 const float** arr, int* ld) {
  arr[ARR_X] = io.x;      ld[ARR_X] = d.dx;
  arr[ARR_S] = io.snode;  ld[ARR_S] = p.ds;
  arr[ARR_E] = io.edata;  ld[ARR_E] = d.de;
  arr[ARR_T] = io.theta;  ld[ARR_T] = d.dtheta;
  arr[ARR_M] = io.mbar;   ld[ARR_M] = p.dm;
  (void)g;
}

int check_io(const ngpde_conv_desc& d, const Plan& p, const ngpde_conv_io& io, bool backward) {
  NGPDE_REQUIRE(io.x != nullptr, "x is NULL");
  NGPDE_REQUIRE(p.ds == 0 || io.snode != nullptr, "snode is NULL but dhs+dpos=%d", p.ds);
  NGPDE_REQUIRE(d.de == 0 || io.edata != nullptr, "edata is NULL but de=%d", d.de);
  NGPDE_REQUIRE(d.dtheta == 0 || d.family != NGPDE_MPPDE_CONV || io.theta != nullptr, "theta is NULL but dtheta=%d",
                d.dtheta);
  NGPDE_REQUIRE(io.phi_params != nullptr, "phi_params is NULL");
  NGPDE_REQUIRE(!p.has_node || io.node_params != nullptr, "node_params is NULL");
  NGPDE_REQUIRE(io.mbar != nullptr, "mbar is NULL");
  NGPDE_REQUIRE(!p.has_node || io.y != nullptr, "y is NULL");
  if (backward) {
    NGPDE_REQUIRE(io.dy != nullptr && io.dx != nullptr && io.dphi_params != nullptr, "backward outputs are NULL");
    NGPDE_REQUIRE(!p.has_node || io.dnode_params != nullptr, "dnode_params is NULL");
  }
  return NGPDE_OK;
}

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }
bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// workspace of the hoisted first layer (forward; the backward recomputes it): folded parameters, the two projections, Q
struct HoistWs {
  size_t off_ft = 0, off_fs = 0, off_fin = 0, off_pt = 0, off_ps = 0, off_q = 0, off_wblk = 0, wblk_bytes = 0;
};
size_t hoist_ws(const Plan& p, int64_t N, size_t off, HoistWs* h) {
  h->off_ft = off;  off = align256(off + sizeof(float) * p.mlp_t.n_params);
  h->off_fs = off;  off = align256(off + sizeof(float) * p.mlp_s.n_params);
  h->off_fin = off; off = align256(off + sizeof(float) * p.phi_in.n_params);
  h->off_pt = h->off_ps = off;  // (the projections are written into Q directly)
  h->off_q = off;   off = align256(off + sizeof(float) * (size_t)N * 2 * p.h_n1);
  h->wblk_bytes = std::max(node_mlp_forward_ws(p.mlp_t), node_mlp_forward_ws(p.mlp_s));
  h->off_wblk = off; off = align256(off + h->wblk_bytes);
  return off;
}

// fold the parameters, project the nodes, interleave: Q = [x|s] Wt + b1 | [x|s] Ws
int hoist_forward(const ngpde_graph* g, const ngpde_conv_desc& d, const Plan& p, const ngpde_conv_io& io, char* ws, const HoistWs& h,
                  cudaStream_t st) {
  float* ft = reinterpret_cast<float*>(ws + h.off_ft);
  float* fs = reinterpret_cast<float*>(ws + h.off_fs);
  float* fin = reinterpret_cast<float*>(ws + h.off_fin);
  float* pt = reinterpret_cast<float*>(ws + h.off_pt);
  float* ps = reinterpret_cast<float*>(ws + h.off_ps);
  float* q = reinterpret_cast<float*>(ws + h.off_q);
  hoist_fold_kernel<<<32, 256, 0, st>>>(io.phi_params, hoist_map(d, p), ft, fs, fin);
  (void)pt; (void)ps;  // the projections land in the two halves of Q's rows directly (strided output of the node kernel)
  if (int rc = node_mlp_forward(g, p.mlp_t, ft, io.x, q, st, ws + h.off_wblk, h.wblk_bytes, io.snode, p.ds, 2 * p.h_n1)) return rc;
  if (int rc = node_mlp_forward(g, p.mlp_s, fs, io.x, q + p.h_n1, st, ws + h.off_wblk, h.wblk_bytes, io.snode, p.ds, 2 * p.h_n1)) return rc;
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

struct NodeHoistWs {
  size_t off_fu = 0, off_fv = 0, off_fin = 0, off_u = 0, off_v = 0, off_q = 0, off_wblk = 0, wblk_bytes = 0;
};
size_t nhoist_ws(const Plan& p, int64_t N, size_t off, NodeHoistWs* h) {
  h->off_fu = off;  off = align256(off + sizeof(float) * p.mlp_u.n_params);
  h->off_fv = off;  off = align256(off + sizeof(float) * p.mlp_v.n_params);
  h->off_fin = off; off = align256(off + sizeof(float) * p.node_in.n_params);
  h->off_u = h->off_v = off;
  h->off_q = off;   off = align256(off + sizeof(float) * (size_t)N * 2 * p.nh_n1);
  h->wblk_bytes = std::max(node_mlp_forward_ws(p.mlp_u), node_mlp_forward_ws(p.mlp_v));
  h->off_wblk = off; off = align256(off + h->wblk_bytes);
  return off;
}
// Q' = [x Wx + b1 | mbar Wm]
int nhoist_forward(const ngpde_graph* g, const ngpde_conv_desc& d, const Plan& p, const ngpde_conv_io& io, char* ws, const NodeHoistWs& h,
                   cudaStream_t st) {
  float* fu = reinterpret_cast<float*>(ws + h.off_fu);
  float* fv = reinterpret_cast<float*>(ws + h.off_fv);
  float* fin = reinterpret_cast<float*>(ws + h.off_fin);
  float* u = reinterpret_cast<float*>(ws + h.off_u);
  float* v = reinterpret_cast<float*>(ws + h.off_v);
  float* q = reinterpret_cast<float*>(ws + h.off_q);
  nhoist_fold_kernel<<<32, 256, 0, st>>>(io.node_params, nhoist_map(d, p), fu, fv, fin);
  (void)u; (void)v;
  if (int rc = node_mlp_forward(g, p.mlp_u, fu, io.x, q, st, ws + h.off_wblk, h.wblk_bytes, nullptr, 0, 2 * p.nh_n1)) return rc;
  if (int rc = node_mlp_forward(g, p.mlp_v, fv, io.mbar, q + p.nh_n1, st, ws + h.off_wblk, h.wblk_bytes, nullptr, 0, 2 * p.nh_n1)) return rc;
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}
// the inner node problem's argument block: x' = Q', input = Q'[:, 0:n1] + Q'[:, n1:2n1]
template <class Args>
void nhoist_args(const Plan& p, const float* q, Args* n) {
  n->arr[ARR_X] = q;
  n->ld[ARR_X] = 2 * p.nh_n1;
  n->n_segs = 2;
  n->segs[0] = Seg{SEG_DST, ARR_X, 0, p.nh_n1, 0};
  n->segs[1] = Seg{SEG_SRC, ARR_X, p.nh_n1, p.nh_n1, 0};
  n->mlp = p.node_in;
}

// ---- the hoisted projections' backward as dense GEMMs over the node axis (tcgen05 3xTF32 engine of the factored GNOConv) ----
constexpr int kWcsRows = 1024;  // rows per block of the weighted column sums
int outer_gemm_splits(int M, int N, int64_t n_nodes, int num_sms) {
  const int tiles = ((M + 127) / 128) * ((N + 63) / 64);
  int sp = std::max(1, (2 * num_sms) / std::max(1, tiles));
  while (sp > 1 && n_nodes / sp < 512) --sp;
  return std::min(sp, 512);
}
// out[M][N] = A' B, A stored [n][M] (lda), B stored [n][N] (ldb): split over the node axis, slices added in order
int node_outer_gemm(const float* A, int lda, int M, const float* B, int ldb, int N, int64_t n_nodes, int num_sms, float* part,
                    float* out, cudaStream_t st) {
  const int sp = outer_gemm_splits(M, N, n_nodes, num_sms);
  if (int rc = gno_gemm(A, lda, true, B, ldb, true, part, N, M, N, n_nodes, sp, nullptr, st)) return rc;
  const int P = M * N;
  reduce_partials_kernel<<<(P + 255) / 256, 256, 0, st>>>(part, sp, P, out);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}
// out[(ds + 1)][C]: rows j < ds = sum_n S[n][j] dq[n][:], row ds = column sums of dq.  C % 4 == 0, C <= 1024, rows 16-byte aligned.
// `part` holds wcs_part_floats(...) floats.
size_t wcs_part_floats(int64_t n_nodes, int C, int ds) {
  const size_t nblk = (size_t)((n_nodes + kWcsRows - 1) / kWcsRows) + 1;
  return (nblk + (nblk + 63) / 64 + 1) * (size_t)(ds + 1) * C;
}
int weighted_colsums(const float* dq, int ld, int C, const float* S, int lds, int ds, int64_t n_nodes, float* part, float* out,
                     cudaStream_t st) {
  const int nblk = std::max(1, (int)((n_nodes + kWcsRows - 1) / kWcsRows));
  const int c4n = C / 4, slices = std::max(1, 256 / c4n);
  const size_t smem = sizeof(float4) * (size_t)slices * (ds + 1) * c4n;
  wcolsum_stage1_kernel<<<nblk, slices * c4n, smem, st>>>(dq, ld, C, S, lds, ds, (long long)n_nodes, kWcsRows, part);
  const int P = (ds + 1) * C;
  const int ngroups = (nblk + 63) / 64;
  float* part2 = part + (size_t)nblk * P;
  reduce_groups_kernel<<<dim3((P + 127) / 128, ngroups), 128, 0, st>>>(part, nblk, P, 64, part2);
  reduce_partials_kernel<<<(P + 255) / 256, 256, 0, st>>>(part2, ngroups, P, out);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

struct BwdLayout {
  int te_e, smem_e, grid_e;
  int te_n, smem_n, grid_n;
  BwdSmem se, sn;
  TcBwdPhase tce, tcn;  // tensor-core variants of the two phases (when eligible)
  size_t off_wt_phi, off_wt_node, off_dmbar, off_dxdirect, off_dxdst, off_desrc, off_part_phi, off_part_node, total;
  // aggr = *: recomputed messages, suffix products, per-edge message cotangents ([E][dm] each) + the forward kernel's tile
  size_t off_msg = 0, off_suf = 0, off_gedge = 0;
  int te_f = 0, smem_f = 0;
  layered::Ws lye, lyn;  // Plan::layered
  size_t off_layered = 0;
  layered::Ws lyg;  // Plan::gno_layered: phi's hidden layers (kept + scratch) and dz [E][K]
  layered::Ws lygn; // Plan::gno_node_gemm: the node update (its cotangent buffers double as dmbar / dx_direct: a region of its own)
  size_t off_dzg = 0;
  // factored GNO
  size_t off_S = 0, off_T = 0, off_DM = 0, off_dBpart = 0, off_B = 0;
  int part_stride = 0, gno_splits = 1;
  // hoisted first layer: recomputed forward part, then dQ and what hangs off it
  NodeHoistWs nhoist;
  size_t off_ndq = 0, off_ng = 0, off_njunk = 0, off_ndfu = 0, off_ndfv = 0, off_ndfin = 0, off_nnodews = 0, nnodews_bytes = 0;
  HoistWs hoist;
  size_t off_dq = 0, off_dpt = 0, off_dps = 0, off_dxt = 0, off_dxs = 0, off_dft = 0, off_dfs = 0, off_dfin = 0, off_nodews = 0,
         nodews_bytes = 0;
  int dxe = 0;  // width of the array the edge phase differentiates: dx, or 2 n1 when hoisted
  // GEMM form of the projections' backward (when the widths allow it)
  bool hgemm = false, nhgemm = false;
  size_t off_fcat = 0, off_dwx = 0, off_dsmall = 0, off_gpart = 0, off_wpart = 0;
};

int bwd_layout(const ngpde_graph* g, const ngpde_conv_desc& d, const Plan& p, BwdLayout* L) {
  L->te_e = 0; L->smem_e = 0; L->grid_e = 0;
  L->dxe = p.hoist ? 2 * p.h_n1 : d.dx;
  if (p.layered) {
    L->te_n = 0; L->smem_n = 0; L->grid_n = 0;
    L->tce = TcBwdPhase{};
    L->tcn = TcBwdPhase{};
    size_t off = 0;
    L->off_wt_phi = L->off_wt_node = L->off_dxdst = L->off_desrc = 0;
    L->off_dmbar = off;     off = align256(off + (p.has_node ? sizeof(float) * g->N * p.dm : 0));
    L->off_dxdirect = off;  off = align256(off + (p.has_node ? sizeof(float) * g->N * d.dx : 0));
    // sized for a call without io.state: kept buffers (recomputed here) followed by the scratch, per phase, sharing the region
    layered::plan_kept(layered::make_phase(p.phi), g->E, off, &L->lye);
    layered::plan_scratch(layered::make_phase(p.phi), g->E, true, true, g->num_sms, L->lye.kept_end, &L->lye);
    if (p.has_node) {
      layered::plan_kept(layered::make_phase(p.node), g->N, off, &L->lyn);
      layered::plan_scratch(layered::make_phase(p.node), g->N, true, true, g->num_sms, L->lyn.kept_end, &L->lyn);
    }
    L->off_layered = off;
    off = std::max(L->lye.end, L->lyn.end);
    L->off_part_phi = L->off_part_node = off;
    L->total = off;
    return NGPDE_OK;
  }
  if (tc_bwd_make(p.hoist ? p.phi_in : p.phi, p.contract, false, false, d.aggr, p.hoist || p.edge_need_dz0, &L->tce)) {
    L->tce.grid = std::max(1, std::min(g->n_units[2], g->num_sms));
    L->grid_e = L->tce.grid;
  } else {
    if (int rc = pick_tile(
            [&](int te) {
              return 4 * bwd_smem(p.phi, p.contract, d.gno_in, d.gno_out, d.aggr, false, p.edge_need_dz0, te, p.gno_Ka).floats;
            },
            &L->te_e, &L->smem_e))
      return rc;
    L->se = bwd_smem(p.phi, p.contract, d.gno_in, d.gno_out, d.aggr, false, p.edge_need_dz0, L->te_e, p.gno_Ka);
    if (int rc = bwd_grid_edge(L->te_e, L->smem_e, std::max(1, g->n_units[tile_index(L->te_e)]), g->num_sms, &L->grid_e))
      return rc;
  }
  L->te_n = 0; L->smem_n = 0; L->grid_n = 0;
  L->tcn = TcBwdPhase{};
  if (p.has_node) {
    if (tc_bwd_make(p.nhoist ? p.node_in : p.node, 0, p.node_addend, true, d.aggr, true, &L->tcn)) {
      L->tcn.grid = std::max(1, std::min((int)((g->N + TC_TILE - 1) / TC_TILE), g->num_sms));
      L->grid_n = L->tcn.grid;
    } else {
      if (int rc = pick_tile([&](int te) { return 4 * bwd_smem(p.node, 0, 0, 0, d.aggr, true, true, te).floats; },
                             &L->te_n, &L->smem_n))
        return rc;
      L->sn = bwd_smem(p.node, 0, 0, 0, d.aggr, true, true, L->te_n);
      const int nu = (int)((g->N + L->te_n - 1) / L->te_n);
      if (int rc = bwd_grid_node(L->te_n, L->smem_n, std::max(1, nu), g->num_sms, &L->grid_n)) return rc;
    }
  }
  size_t off = 0;
  if (L->tce.on) { L->tce.ws_off = off; off = align256(off + 4 * (size_t)L->tce.lay.block_floats); }
  if (L->tcn.on) { L->tcn.ws_off = off; off = align256(off + 4 * (size_t)L->tcn.lay.block_floats); }
  L->off_wt_phi = off;    off = align256(off + sizeof(float) * p.phi.n_params);
  L->off_wt_node = off;   off = align256(off + sizeof(float) * p.node.n_params);
  L->off_dmbar = off;     off = align256(off + (p.has_node ? sizeof(float) * g->N * p.dm : 0));
  L->off_dxdirect = off;  off = align256(off + (p.has_node ? sizeof(float) * g->N * d.dx : 0));
  L->off_dxdst = off;     off = align256(off + ((p.edge_dst_side || p.hoist) ? sizeof(float) * g->N * L->dxe : 0));
  // per-edge source-side cotangents: [E][dx], or [E][n1] for a hoisted first layer (only Q's source half has a source side)
  L->off_desrc = off;     off = align256(off + sizeof(float) * g->E * (p.hoist ? p.h_n1 : d.dx));
  L->part_stride = p.hoist ? p.phi_in.n_params : p.phi.n_params;
  if (d.aggr == NGPDE_AGGR_PROD) {
    NGPDE_REQUIRE(!L->tce.on && !p.hoist && p.contract != 2, "internal: aggr = * runs on the FFMA kernels");
    if (int rc = pick_tile([&](int t) { return 4 * fwd_smem(p.phi, p.contract, d.gno_in, d.gno_out, t, p.gno_Ka).floats; },
                           &L->te_f, &L->smem_f))
      return rc;
    L->off_msg = off;   off = align256(off + sizeof(float) * g->E * p.dm);
    L->off_suf = off;   off = align256(off + sizeof(float) * g->E * p.dm);
    L->off_gedge = off; off = align256(off + sizeof(float) * g->E * p.dm);
  }
  if (p.nhoist) {
    NGPDE_REQUIRE(L->tcn.on, "internal: hoisted node plan without a tensor-core node phase");
    off = nhoist_ws(p, g->N, off, &L->nhoist);
    L->off_ndq = off;   off = align256(off + sizeof(float) * g->N * 2 * p.nh_n1);
    L->off_ng = L->off_njunk = off;
    L->off_ndfu = off;  off = align256(off + sizeof(float) * p.mlp_u.n_params);
    L->off_ndfv = off;  off = align256(off + sizeof(float) * p.mlp_v.n_params);
    L->off_ndfin = off; off = align256(off + sizeof(float) * p.node_in.n_params);
    L->nnodews_bytes = std::max(node_mlp_backward_ws(g, p.mlp_u), node_mlp_backward_ws(g, p.mlp_v));
    L->off_nnodews = off; off = align256(off + L->nnodews_bytes);
  }
  if (p.hoist) {
    NGPDE_REQUIRE(L->tce.on, "internal: hoisted plan without a tensor-core edge phase");
    off = hoist_ws(p, g->N, off, &L->hoist);
    L->off_dq = off;   off = align256(off + sizeof(float) * g->N * 2 * p.h_n1);
    L->off_dpt = L->off_dps = off;
    L->off_dxt = off;  off = align256(off + sizeof(float) * g->N * d.dx);
    L->off_dxs = off;  off = align256(off + sizeof(float) * g->N * d.dx);
    L->off_dft = off;  off = align256(off + sizeof(float) * p.mlp_t.n_params);
    L->off_dfs = off;  off = align256(off + sizeof(float) * p.mlp_s.n_params);
    L->off_dfin = off; off = align256(off + sizeof(float) * p.phi_in.n_params);
    L->nodews_bytes = std::max(node_mlp_backward_ws(g, p.mlp_t), node_mlp_backward_ws(g, p.mlp_s));
    L->off_nodews = off; off = align256(off + L->nodews_bytes);
  }
  if (p.contract == 2) {
    const size_t R = (size_t)p.gno_Ka * d.gno_in;
    L->part_stride = p.phi.w_off[p.phi.L - 1];  // the last layer's gradient comes from the GEMM dB = S' DM
    // split-K slices of dB = S' DM: m-tiles x slices should fill whole waves of 2 CTAs per SM
    {
      const int mt = (int)((R + 127) / 128);
      const int per_wave = 2 * g->num_sms;
      int best = 1;
      double best_cost = 1e30;
      for (int sp = 1; sp <= 48; ++sp) {
        if ((int64_t)sp * 1024 > g->N && sp > 1) break;
        const int waves = (mt * sp + per_wave - 1) / per_wave;
        const double cost = (double)waves / sp;  // time ~ waves x (K / slices)
        if (cost < best_cost - 1e-12) { best_cost = cost; best = sp; }
      }
      L->gno_splits = best;
    }
    L->off_S = off;      off = align256(off + sizeof(float) * (size_t)g->N * R);
    L->off_T = off;      off = align256(off + sizeof(float) * (size_t)g->N * R);
    L->off_DM = off;     off = align256(off + sizeof(float) * (size_t)g->N * d.gno_out);
    L->off_dBpart = off; off = align256(off + sizeof(float) * (size_t)L->gno_splits * R * d.gno_out);
    L->off_B = off;      off = align256(off + sizeof(float) * R * d.gno_out);
    if (p.gno_layered) {
      const layered::Phase hp = layered::make_phase(p.phi_hidden);
      layered::plan_kept(hp, g->E, off, &L->lyg);
      layered::plan_scratch(hp, g->E, true, true, g->num_sms, L->lyg.kept_end, &L->lyg);
      off = L->lyg.end;
      L->off_dzg = off;  off = align256(off + sizeof(float) * (size_t)g->E * p.gno_K);
      if (p.gno_node_gemm) {
        const layered::Phase np = layered::make_phase(p.node);
        layered::plan_kept(np, g->N, off, &L->lygn);
        layered::plan_scratch(np, g->N, true, true, g->num_sms, L->lygn.kept_end, &L->lygn);
        off = L->lygn.end;
      }
    }
  }
  if (p.hoist || p.nhoist) {
    L->hgemm = p.hoist && (d.dx & 3) == 0 && p.ds <= 7;
    L->nhgemm = p.nhoist && (d.dx & 3) == 0 && (p.dm & 3) == 0;
    const int n1m = std::max(p.hoist ? p.h_n1 : 0, p.nhoist ? p.nh_n1 : 0);
    const int mm = std::max(d.dx, p.dm);
    L->off_fcat = off;   off = align256(off + sizeof(float) * (size_t)(d.dx + p.ds) * 2 * n1m);
    L->off_dwx = off;    off = align256(off + sizeof(float) * (size_t)mm * 2 * n1m);
    L->off_dsmall = off; off = align256(off + sizeof(float) * 8 * 2 * n1m);
    L->off_gpart = off;  off = align256(off + sizeof(float) * (size_t)outer_gemm_splits(mm, 2 * n1m, g->N, g->num_sms) * mm * 2 * n1m * 2);
    L->off_wpart = off;  off = align256(off + sizeof(float) * wcs_part_floats(g->N, 2 * n1m, 7));
  }
  L->off_part_phi = off;  off = align256(off + sizeof(float) * (size_t)L->grid_e * L->part_stride);
  L->off_part_node = off; off = align256(off + sizeof(float) * (size_t)L->grid_n * (p.nhoist ? p.node_in.n_params : p.node.n_params));
  L->total = off;
  return NGPDE_OK;
}


}  // namespace

// ---- a bare Chain of Dense layers over the node axis (used by GCNConv's W*x) ----
int make_mlp_dev(const ngpde_mlp& m, MlpDev* out, const char* what) { return make_mlp(m, out, what); }

namespace {
bool node_tc_fwd_plan(const MlpDev& mlp, TcPhase* t) {
  *t = TcPhase{};
  t->on = tc_make_layout(mlp, 0, false, mlp.dims[mlp.L], true, &t->lay, &t->smem, &t->off_cols, &t->off_groups, &t->group_bytes);
  return t->on;
}
}  // namespace

// bytes of the prepared weight block the tensor-core node kernel wants (0: the MLP runs on the FFMA engine)
size_t node_mlp_forward_ws(const MlpDev& mlp) {
  TcPhase t;
  return node_tc_fwd_plan(mlp, &t) ? align256(4 * (size_t)t.lay.block_floats) : 0;
}

namespace {
// node input = [x (dx columns) ; snode (ds columns)]; ds = 0: x alone
template <class Args>
void node_input_segs(Args* n, const float* x, int dx, const float* snode, int ds) {
  n->arr[ARR_X] = x;
  n->ld[ARR_X] = dx;
  n->n_segs = 1;
  n->segs[0] = Seg{SEG_DST, ARR_X, 0, dx, 0};
  if (ds > 0) {
    n->arr[ARR_S] = snode;
    n->ld[ARR_S] = ds;
    n->segs[1] = Seg{SEG_DST, ARR_S, 0, ds, dx};
    n->n_segs = 2;
  }
}
}  // namespace

int node_mlp_forward(const ngpde_graph* g, const MlpDev& mlp, const float* params, const float* x, float* y,
                     cudaStream_t st, void* workspace, size_t ws_bytes, const float* snode, int ds, int out_ld) {
  FwdArgs n{};
  n.out_ld = out_ld;
  node_input_segs(&n, x, mlp.dims[0] - ds, snode, ds);
  n.mlp = mlp;
  n.params = params;
  n.dout = mlp.dims[mlp.L];
  n.out = y;
  TcPhase t;
  if (workspace != nullptr && aligned16(workspace) && node_tc_fwd_plan(mlp, &t) && ws_bytes >= 4 * (size_t)t.lay.block_floats) {
    // tcgen05 node kernel (ngpde_tc.cuh): layers <= 64 wide
    n.tg = TileGraph{nullptr, nullptr, nullptr, nullptr, nullptr, (int)((g->N + TC_TILE - 1) / TC_TILE), (int)g->N, 1};
    return launch_fwd_tc(true, g->num_sms, t, mlp, params, n, static_cast<float*>(workspace), st);
  }
  NGPDE_REQUIRE(out_ld == 0 || out_ld == n.dout, "internal: strided node-MLP output needs the tensor-core kernel");
  int te = 0, smem = 0;
  if (int rc = pick_tile([&](int t) { return 4 * fwd_smem(mlp, 0, 0, 0, t).floats; }, &te, &smem)) return rc;
  FwdSmem fs = fwd_smem(mlp, 0, 0, 0, te);
  n.tg = TileGraph{nullptr, nullptr, nullptr, nullptr, nullptr, (int)((g->N + te - 1) / te), (int)g->N, 1};
  n.offA = fs.offA; n.offB = fs.offB; n.offW = fs.offW; n.offH = fs.offH;
  return launch_fwd_node(te, n, smem, g->num_sms, st);
}

namespace {
struct NodeBwdLayout {
  int te, smem, grid;
  BwdSmem s;
  TcBwdPhase tc;
  size_t off_wt, off_part, total;
};
int node_bwd_layout(const ngpde_graph* g, const MlpDev& mlp, NodeBwdLayout* L) {
  size_t off = 0;
  L->tc = TcBwdPhase{};
  if (tc_bwd_make(mlp, 0, false, true, NGPDE_AGGR_SUM, true, &L->tc)) {  // tcgen05 node kernel (ngpde_tc_bwd.cuh)
    L->te = TC_TILE;
    L->smem = L->tc.smem;
    L->tc.grid = std::max(1, std::min((int)((g->N + TC_TILE - 1) / TC_TILE), g->num_sms));
    L->grid = L->tc.grid;
    L->tc.ws_off = off;
    off = align256(off + 4 * (size_t)L->tc.lay.block_floats);
  } else {
    if (int rc = pick_tile([&](int te) { return 4 * bwd_smem(mlp, 0, 0, 0, 0, true, true, te).floats; }, &L->te, &L->smem))
      return rc;
    L->s = bwd_smem(mlp, 0, 0, 0, 0, true, true, L->te);
    const int nu = (int)((g->N + L->te - 1) / L->te);
    if (int rc = bwd_grid_node(L->te, L->smem, std::max(1, nu), g->num_sms, &L->grid)) return rc;
  }
  L->off_wt = off;   off = align256(off + sizeof(float) * mlp.n_params);
  L->off_part = off; off = align256(off + sizeof(float) * (size_t)L->grid * mlp.n_params);
  L->total = off;
  return NGPDE_OK;
}
}  // namespace

bool node_mlp_backward_fuses_act(const ngpde_graph* g, const MlpDev& mlp) {
  NodeBwdLayout L;
  return node_bwd_layout(g, mlp, &L) == NGPDE_OK && L.tc.on;
}

size_t node_mlp_backward_ws(const ngpde_graph* g, const MlpDev& mlp) {
  NodeBwdLayout L;
  if (node_bwd_layout(g, mlp, &L)) return 0;
  return L.total;
}

// dy -> dx (may be nullptr... it is always produced here), dparams
int node_mlp_backward(const ngpde_graph* g, const MlpDev& mlp, const float* params, const float* x, const float* dy,
                      float* dx, float* dparams, void* workspace, size_t ws_bytes, cudaStream_t st, const float* snode, int ds,
                      int dy_ld, const float* yact, int yact_kind) {
  NodeBwdLayout L;
  if (int rc = node_bwd_layout(g, mlp, &L)) return rc;
  if (ws_bytes < L.total) {
    set_error("node MLP workspace too small: %zu < %zu", ws_bytes, L.total);
    return NGPDE_ERR_WORKSPACE;
  }
  char* ws = static_cast<char*>(workspace);
  float* wt = reinterpret_cast<float*>(ws + L.off_wt);
  float* part = reinterpret_cast<float*>(ws + L.off_part);
  NGPDE_CUDA_TRY(cudaMemsetAsync(part, 0, L.total - L.off_part, st));
  if (!L.tc.on) transpose_weights_kernel<<<64, 256, 0, st>>>(params, wt, mlp);
  BwdArgs n{};
  n.tg = TileGraph{nullptr, nullptr, nullptr, nullptr, nullptr, (int)((g->N + L.te - 1) / L.te), (int)g->N, 1};
  node_input_segs(&n, x, mlp.dims[0] - ds, snode, ds);
  n.mlp = mlp;
  n.params = params;
  n.wt = wt;
  n.dout = mlp.dims[mlp.L];
  n.gout_ptr = dy;
  n.dparams_partial = part;
  n.part_stride = mlp.n_params;
  n.dx_direct = dx;
  n.dx = mlp.dims[0] - ds;
  n.need_dz0 = 1;
  n.store_last = L.s.store_last;
  std::memcpy(n.zoff, L.s.zoff, sizeof(n.zoff));
  n.offG0 = L.s.offG0; n.offG1 = L.s.offG1; n.offW = L.s.offW;
  n.gout_ld = dy_ld;
  NGPDE_REQUIRE(yact == nullptr || L.tc.on, "internal: fused act'(y) needs the tensor-core node kernel");
  n.yact = yact;
  n.yact_kind = yact_kind;
  if (L.tc.on) {
    if (int rc = launch_bwd_tc(true, L.tc, mlp, n, reinterpret_cast<float*>(ws + L.tc.ws_off), st)) return rc;
  } else {
    NGPDE_REQUIRE(dy_ld == 0 || dy_ld == n.dout, "internal: strided node-MLP cotangent needs the tensor-core kernel");
    if (int rc = launch_bwd_node(L.te, n, L.smem, L.grid, st)) return rc;
  }
  reduce_partials_kernel<<<(mlp.n_params + 255) / 256, 256, 0, st>>>(part, L.grid, mlp.n_params, dparams);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

}  // namespace ngpde

using namespace ngpde;

namespace {

struct FwdPlan {
  NodeHoistWs nhoist;
  TcPhase edge, node;
  size_t ws_bytes = 0;
  size_t off_S = 0, off_B = 0;
  HoistWs hoist;
  layered::Ws lye, lyn;  // Plan::layered
  layered::Ws lyg, lygn;  // Plan::gno_layered (edge phase: phi's hidden layers; node update)
};

// layout of the optional ngpde_conv_io.state buffer: the hoisted projections and folded parameters, kept for the backward
struct StatePlan {
  HoistWs hoist;
  NodeHoistWs nhoist;
  layered::Ws lye, lyn;  // Plan::layered: the kept activations of both phases (up to 8 GiB; beyond that the backward recomputes)
  layered::Ws lyg;       // Plan::gno_layered: phi's hidden activations, z = the activated last hidden layer [E][K] and the
  size_t off_zg = 0;     // per-destination sums S [N][Ka * gin] (up to 40 GiB; beyond that the backward recomputes)
  size_t off_Sg = 0;
  size_t bytes = 0;
};
StatePlan state_plan(const Plan& p, int64_t N, int64_t E = 0) {
  StatePlan sp;
  size_t off = 0;
  if (p.layered) {
    layered::plan_kept(layered::make_phase(p.phi), E, 0, &sp.lye);
    off = sp.lye.kept_end;
    if (p.has_node) {
      layered::plan_kept(layered::make_phase(p.node), N, off, &sp.lyn);
      off = sp.lyn.kept_end;
    }
    sp.bytes = off <= (size_t(8) << 30) ? off : 0;
    return sp;
  }
  if (p.gno_layered) {
    layered::plan_kept(layered::make_phase(p.phi_hidden), E, 0, &sp.lyg);
    sp.off_zg = sp.lyg.kept_end;
    off = align256(sp.off_zg + sizeof(float) * (size_t)E * p.gno_K);
    sp.off_Sg = off;
    const int gin = p.phi.dims[p.phi.L] / std::max(1, p.dm);  // phi's last layer is gin * gout wide
    off = align256(off + sizeof(float) * (size_t)N * p.gno_Ka * gin);
    sp.bytes = off <= (size_t(40) << 30) ? off : 0;
    return sp;
  }
  if (p.hoist) off = hoist_ws(p, N, off, &sp.hoist);
  if (p.nhoist) off = nhoist_ws(p, N, off, &sp.nhoist);
  sp.bytes = off;
  return sp;
}

FwdPlan fwd_plan(const Plan& p, int aggr, int64_t N, int gin, int64_t E = 0) {
  FwdPlan f;
  size_t off = 0;
  if (p.layered) {  // the two phases run one after the other: they share the region (sized for a call without io.state)
    layered::plan_scratch(layered::make_phase(p.phi), E, false, false, 1, 0, &f.lye);
    if (p.has_node) layered::plan_scratch(layered::make_phase(p.node), N, false, false, 1, 0, &f.lyn);
    f.ws_bytes = std::max(f.lye.end, f.lyn.end);
    return f;
  }
  if (p.hoist) off = hoist_ws(p, N, off, &f.hoist);
  const MlpDev& ephi = p.hoist ? p.phi_in : p.phi;
  // max/min: the backward's tie mask compares recomputed messages with the forward's bit for bit, so both must run
  // the same arithmetic -- those aggregations stay on the FFMA kernels until the backward has a tensor-core twin.
  const bool aggr_ok = aggr == NGPDE_AGGR_SUM || aggr == NGPDE_AGGR_MEAN;
  f.edge.on = aggr_ok && tc_make_layout(ephi, p.contract, false, p.dm, false, &f.edge.lay, &f.edge.smem, &f.edge.off_cols,
                                        &f.edge.off_groups, &f.edge.group_bytes);
  if (f.edge.on) {
    f.edge.ws_off = off;
    off = align256(off + 4 * (size_t)f.edge.lay.block_floats);
  }
  if (p.nhoist) off = nhoist_ws(p, N, off, &f.nhoist);
  if (p.has_node) {
    f.node.on = tc_make_layout(p.nhoist ? p.node_in : p.node, 0, p.node_addend, p.dy, true, &f.node.lay, &f.node.smem, &f.node.off_cols,
                               &f.node.off_groups, &f.node.group_bytes);
    if (f.node.on) {
      f.node.ws_off = off;
      off = align256(off + 4 * (size_t)f.node.lay.block_floats);
    }
  }
  if (p.contract == 2) {
    f.off_S = off;
    off = align256(off + sizeof(float) * (size_t)N * p.gno_Ka * gin);
    f.off_B = off;  // 16-byte aligned copy of B = [W3; b3] for when the flat parameter segment is not
    off = align256(off + sizeof(float) * (size_t)p.gno_Ka * gin * p.dm);
    if (p.gno_layered) {
      layered::plan_scratch(layered::make_phase(p.phi_hidden), E, false, false, 1, off, &f.lyg);
      off = f.lyg.end;
      if (p.gno_node_gemm) {
        layered::plan_scratch(layered::make_phase(p.node), N, false, false, 1, off, &f.lygn);
        off = f.lygn.end;
      }
    }
  }
  f.ws_bytes = off;
  return f;
}


// ---- Plan::layered: every Dense layer as one tcgen05 GEMM over all edges / nodes (ngpde_layered.cuh) ----
layered::Gather layered_gather(const ngpde_graph* g, const ngpde_conv_desc& d, const Plan& p, const ngpde_conv_io& io, bool node) {
  layered::Gather ga{};
  fill_arrays(g, d, p, io, ga.arr, ga.ld);
  ga.n_segs = node ? p.n_nsegs : p.n_esegs;
  std::memcpy(ga.segs, node ? p.nsegs : p.esegs, sizeof(p.esegs));
  ga.src = node ? nullptr : g->src;
  ga.dst = node ? nullptr : g->dst;
  ga.perm = node ? nullptr : g->perm;
  ga.gdiv = (int)std::max<int64_t>(1, (node ? g->N : g->E) / std::max<int64_t>(1, g->G));
  return ga;
}

int layered_forward(const ngpde_graph* g, const ngpde_conv_desc& d, const Plan& p, const ngpde_conv_io& io, const FwdPlan& fp,
                    char* ws, cudaStream_t st) {
  NGPDE_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "workspace must be 16-byte aligned");
  NGPDE_REQUIRE(aligned16(io.mbar) && (!p.has_node || aligned16(io.y)), "layered evaluation: mbar and y must be 16-byte aligned");
  // with io.state the activations the backward needs are kept there (ngpde_conv_state_bytes), else nothing is kept
  const StatePlan sp = state_plan(p, g->N, g->E);
  const bool keep = io.state != nullptr && sp.bytes > 0;
  NGPDE_REQUIRE(!keep || aligned16(io.state), "io.state must be 16-byte aligned");
  char* kbase = keep ? static_cast<char*>(io.state) : ws;
  {
    ProfScope prof(NGPDE_PROF_FWD_EDGE, st);
    const layered::Phase ph = layered::make_phase(p.phi);
    layered::Ws w = keep ? sp.lye : fp.lye;
    if (keep) layered::plan_scratch(ph, g->E, false, true, 1, 0, &w);
    const float* msg = nullptr;
    if (int rc = layered::run_forward(ph, w, kbase, ws, layered_gather(g, d, p, io, false), g->E, io.phi_params, keep, nullptr, &msg, st))
      return rc;
    const long long tot = (long long)g->N * (p.dm / 4);
    layered::aggregate_rows_kernel<<<layered::blocks(tot, 256), 256, 0, st>>>((int)g->N, p.dm, d.aggr == NGPDE_AGGR_MEAN, g->rowptr,
                                                                              msg, io.mbar);
  }
  if (p.has_node) {
    ProfScope prof(NGPDE_PROF_FWD_NODE, st);
    const layered::Phase ph = layered::make_phase(p.node);
    layered::Ws w = keep ? sp.lyn : fp.lyn;
    if (keep) layered::plan_scratch(ph, g->N, false, true, 1, 0, &w);
    if (int rc = layered::run_forward(ph, w, kbase, ws, layered_gather(g, d, p, io, true), g->N, io.node_params, keep, io.y, nullptr, st))
      return rc;
  }
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

int layered_backward(const ngpde_graph* g, const ngpde_conv_desc& d, const Plan& p, const ngpde_conv_io& io, const BwdLayout& L,
                     char* ws, cudaStream_t st) {
  float* dmbar = p.has_node ? reinterpret_cast<float*>(ws + L.off_dmbar) : nullptr;
  float* dxdirect = p.has_node ? reinterpret_cast<float*>(ws + L.off_dxdirect) : nullptr;
  NGPDE_REQUIRE(aligned16(io.dy), "layered evaluation: dy must be 16-byte aligned");
  // the forward's activations: left in io.state by the forward call, or recomputed here into the workspace
  const StatePlan sp = state_plan(p, g->N, g->E);
  const bool kept = io.state != nullptr && sp.bytes > 0;
  NGPDE_REQUIRE(!kept || aligned16(io.state), "io.state must be 16-byte aligned");
  char* kbase = kept ? static_cast<char*>(io.state) : ws;
  if (p.has_node) {
    ProfScope prof(NGPDE_PROF_BWD_NODE, st);
    const layered::Phase ph = layered::make_phase(p.node);
    const layered::Gather ga = layered_gather(g, d, p, io, true);
    layered::Ws w = L.lyn;
    if (kept) { w = sp.lyn; layered::plan_scratch(ph, g->N, true, true, g->num_sms, L.off_layered, &w); }
    else if (int rc = layered::run_forward(ph, w, kbase, ws, ga, g->N, io.node_params, true, nullptr, nullptr, st)) return rc;
    const float* dz0 = nullptr;
    if (int rc = layered::run_backward(ph, w, kbase, ws, g->N, io.dy, nullptr, nullptr, true, g->num_sms, io.dnode_params, &dz0, st))
      return rc;
    const long long tot = (long long)g->N * (d.dx + p.dm);
    layered::node_split_kernel<<<layered::blocks(tot, 256), 256, 0, st>>>(ga, (int)g->N, d.dx, p.dm, ph.ld[0], dz0, dxdirect, dmbar);
  }
  {
    ProfScope prof(NGPDE_PROF_BWD_EDGE, st);
    const layered::Phase ph = layered::make_phase(p.phi);
    const layered::Gather ga = layered_gather(g, d, p, io, false);
    layered::Ws w = L.lye;
    if (kept) { w = sp.lye; layered::plan_scratch(ph, g->E, true, true, g->num_sms, L.off_layered, &w); }
    else if (int rc = layered::run_forward(ph, w, kbase, ws, ga, g->E, io.phi_params, true, nullptr, nullptr, st)) return rc;
    const float* dz0 = nullptr;
    if (int rc = layered::run_backward(ph, w, kbase, ws, g->E, p.has_node ? dmbar : io.dy, g->dst,
                                       d.aggr == NGPDE_AGGR_MEAN ? g->rowptr : nullptr, p.edge_need_dz0, g->num_sms,
                                       io.dphi_params, &dz0, st))
      return rc;
    const long long tot = (long long)g->N * d.dx;
    if (p.edge_need_dz0) {
      if (layered::quads_ok(ga, d.dx) && aligned16(io.dx))
        layered::edge_dx4_kernel<<<layered::blocks(tot / 4, 256), 256, 0, st>>>(ga, (int)g->N, d.dx, ph.ld[0], g->rowptr, g->tptr, g->tpos,
                                                                                dz0, dxdirect, io.dx);
      else
        layered::edge_dx_kernel<<<layered::blocks(tot, 256), 256, 0, st>>>(ga, (int)g->N, d.dx, ph.ld[0], g->rowptr, g->tptr, g->tpos, dz0,
                                                                           dxdirect, io.dx);
    } else if (dxdirect) {
      NGPDE_CUDA_TRY(cudaMemcpyAsync(io.dx, dxdirect, sizeof(float) * tot, cudaMemcpyDeviceToDevice, st));
    } else {
      NGPDE_CUDA_TRY(cudaMemsetAsync(io.dx, 0, sizeof(float) * tot, st));
    }
  }
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

}  // namespace

extern "C" int ngpde_set_option(int32_t option, int32_t value) {
  switch (option) {
    case NGPDE_OPT_TENSOR_CORES: tc_set_enabled(value != 0); return NGPDE_OK;
    case NGPDE_OPT_GNO_FACTORED: g_gno_factored = value != 0; return NGPDE_OK;
    case NGPDE_OPT_DEBUG_SKIP: g_debug_skip = value; return NGPDE_OK;
    case NGPDE_OPT_HOIST: g_hoist = value != 0; return NGPDE_OK;
    case NGPDE_OPT_LAYERED: g_layered = value != 0; return NGPDE_OK;
    case NGPDE_OPT_GNO_LAYERED: g_gno_layered = value != 0; return NGPDE_OK;
    default: set_error("unknown option %d", option); return NGPDE_ERR_INVALID;
  }
}

extern "C" size_t ngpde_conv_workspace_bytes(ngpde_graph_t g, const ngpde_conv_desc* desc, int32_t backward) {
  if (!g || !desc) return 0;
  Plan p;
  if (make_plan(g, *desc, &p)) return 0;
  if (!backward) return fwd_plan(p, desc->aggr, g->N, desc->gno_in, g->E).ws_bytes + 256;
  BwdLayout L;
  if (bwd_layout(g, *desc, p, &L)) return 0;
  return L.total + 256;
}

extern "C" size_t ngpde_conv_state_bytes(ngpde_graph_t g, const ngpde_conv_desc* desc) {
  if (!g || !desc) return 0;
  Plan p;
  if (make_plan(g, *desc, &p)) return 0;
  const size_t b = state_plan(p, g->N, g->E).bytes;
  return b ? b + 256 : 0;
}

extern "C" int ngpde_conv_kernel_paths(ngpde_graph_t g, const ngpde_conv_desc* desc, int32_t* paths) {
  NGPDE_REQUIRE(g && desc && paths, "null argument");
  Plan p;
  if (int rc = make_plan(g, *desc, &p)) return rc;
  const FwdPlan fp = fwd_plan(p, desc->aggr, g->N, desc->gno_in, g->E);
  BwdLayout L;
  if (int rc = bwd_layout(g, *desc, p, &L)) return rc;
  if (p.layered) {  // 3: one tcgen05 GEMM per Dense layer (ngpde_layered.cuh)
    paths[NGPDE_PROF_FWD_EDGE] = paths[NGPDE_PROF_BWD_EDGE] = 3;
    paths[NGPDE_PROF_FWD_NODE] = paths[NGPDE_PROF_BWD_NODE] = p.has_node ? 3 : -1;
    return NGPDE_OK;
  }
  paths[NGPDE_PROF_FWD_EDGE] = fp.edge.on ? 1 : (p.contract == 2 ? 2 : 0);
  paths[NGPDE_PROF_FWD_NODE] = p.has_node ? (fp.node.on ? 1 : 0) : -1;
  paths[NGPDE_PROF_BWD_EDGE] = L.tce.on ? 1 : (p.contract == 2 ? 2 : 0);
  paths[NGPDE_PROF_BWD_NODE] = p.has_node ? (L.tcn.on ? 1 : 0) : -1;
  if (p.gno_node_gemm) paths[NGPDE_PROF_FWD_NODE] = paths[NGPDE_PROF_BWD_NODE] = 3;
  return NGPDE_OK;
}

extern "C" int ngpde_conv_forward(ngpde_graph_t g, const ngpde_conv_desc* desc, const ngpde_conv_io* io,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  NGPDE_REQUIRE(g && desc && io, "null argument");
  Plan p;
  if (int rc = make_plan(g, *desc, &p)) return rc;
  if (g->N == 0) return NGPDE_OK;
  if (int rc = check_io(*desc, p, *io, false)) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const FwdPlan fp = fwd_plan(p, desc->aggr, g->N, desc->gno_in, g->E);
  if (fp.ws_bytes > 0 && (workspace == nullptr || workspace_bytes < fp.ws_bytes)) {
    set_error("forward workspace too small: %zu bytes given, %zu needed (ngpde_conv_workspace_bytes(g, desc, 0))",
              workspace_bytes, fp.ws_bytes);
    return NGPDE_ERR_WORKSPACE;
  }
  char* fws = static_cast<char*>(workspace);
  // the hoisted projections live in io->state when the caller provides it (the backward then reuses them), else in the workspace
  const StatePlan sp = state_plan(p, g->N, g->E);
  const bool keep = io->state != nullptr && sp.bytes > 0;
  NGPDE_REQUIRE(!keep || aligned16(io->state), "io.state must be 16-byte aligned");
  char* hbase = keep ? static_cast<char*>(io->state) : fws;
  const HoistWs& hw = keep ? sp.hoist : fp.hoist;
  const NodeHoistWs& nhw = keep ? sp.nhoist : fp.nhoist;

  if (p.layered) return layered_forward(g, *desc, p, *io, fp, fws, st);

  // ---- edge phase ----
  int te = 0, smem = 0;
  if (int rc = pick_tile(
          [&](int t) { return 4 * fwd_smem(p.phi, p.contract, desc->gno_in, desc->gno_out, t, p.gno_Ka).floats; }, &te, &smem))
    return rc;
  FwdSmem fs = fwd_smem(p.phi, p.contract, desc->gno_in, desc->gno_out, te, p.gno_Ka);
  FwdArgs a{};
  const int ti = tile_index(te);
  a.tg = TileGraph{g->rowptr, g->src, g->dst, g->perm, g->units[ti], g->n_units[ti], (int)g->N,
                   (int)std::max<int64_t>(1, g->E / std::max<int64_t>(1, g->G))};
  fill_arrays(g, *desc, p, *io, a.arr, a.ld);
  a.n_segs = p.n_esegs;
  std::memcpy(a.segs, p.esegs, sizeof(p.esegs));
  a.mlp = p.phi;
  a.params = io->phi_params;
  a.contract = p.contract;
  a.gin = desc->gno_in;
  a.gout = desc->gno_out;
  a.aggr = desc->aggr;
  a.dout = p.dm;
  a.out = io->mbar;
  a.addend = nullptr;
  a.offA = fs.offA; a.offB = fs.offB; a.offW = fs.offW; a.offH = fs.offH; a.offZt = fs.offZt;
  a.gno_Ka = p.gno_Ka;
  a.gno_S = p.contract == 2 ? reinterpret_cast<float*>(fws + fp.off_S) : nullptr;
  if (p.gno_layered && keep) a.gno_S = reinterpret_cast<float*>(static_cast<char*>(io->state) + sp.off_Sg);  // kept for dB = S' DM
  {
    ProfScope prof(NGPDE_PROF_FWD_EDGE, st);
    if (p.hoist) {
      // first layer hoisted to the nodes; the edge kernel runs the inner problem on Q (Plan::hoist)
      NGPDE_REQUIRE(fp.edge.on, "internal: hoisted plan without a tensor-core edge phase");
      if (int rc = hoist_forward(g, *desc, p, *io, hbase, hw, st)) return rc;
      a.arr[ARR_X] = reinterpret_cast<const float*>(hbase + hw.off_q);
      a.ld[ARR_X] = 2 * p.h_n1;
      a.n_segs = 2;
      a.segs[0] = p.hsegs[0];
      a.segs[1] = p.hsegs[1];
      a.mlp = p.phi_in;
      a.params = reinterpret_cast<const float*>(hbase + hw.off_fin);
      a.skip_l0 = 1;
      a.tg.unit_ptr = g->units[2];
      a.tg.n_units = g->n_units[2];
      if (int rc = launch_fwd_tc(false, g->num_sms, fp.edge, p.phi_in, a.params, a, reinterpret_cast<float*>(fws + fp.edge.ws_off), st))
        return rc;
    } else if (fp.edge.on) {
      a.tg.unit_ptr = g->units[2];
      a.tg.n_units = g->n_units[2];
      if (int rc = launch_fwd_tc(false, g->num_sms, fp.edge, p.phi, io->phi_params, a, reinterpret_cast<float*>(fws + fp.edge.ws_off), st))
        return rc;
    } else if (p.gno_layered) {
      // z_e = phi's hidden layers as GEMMs over all edges, then S_n = sum_e [z_e; 1] h_e' with one warp per destination
      const layered::Phase hp = layered::make_phase(p.phi_hidden);
      const float* z = nullptr;
      if (keep) {  // hidden activations and z stay in io->state for the backward
        if (int rc = layered::run_forward(hp, sp.lyg, static_cast<char*>(io->state), fws, layered_gather(g, *desc, p, *io, false), g->E,
                                          io->phi_params, true, reinterpret_cast<float*>(static_cast<char*>(io->state) + sp.off_zg), &z, st))
          return rc;
      } else if (int rc = layered::run_forward(hp, fp.lyg, fws, fws, layered_gather(g, *desc, p, *io, false), g->E, io->phi_params,
                                               false, nullptr, &z, st)) {
        return rc;
      }
      gnonode::Args na{};
      na.z = z; na.x = io->x; na.ldx = desc->dx; na.src = g->src; na.rowptr = g->rowptr; na.N = (int)g->N; na.Ka = p.gno_Ka;
      na.S = a.gno_S;
      if (int rc = gnonode::launch(na, p.gno_K, false, g->num_sms, st)) return rc;
    } else {
      if (int rc = launch_fwd_edge(te, a, smem, g->num_sms, st)) return rc;
    }
    if (p.contract == 2) {
      // mbar = (S B) ./ deg,  B = [W3; b3] = the last layer's flat parameter segment viewed as [Ka*gin][gout]
      const int R = p.gno_Ka * desc->gno_in;
      NGPDE_REQUIRE(aligned16(io->x) && aligned16(io->mbar),
                    "GNOConv (factored evaluation): x and mbar must be 16-byte aligned (or set NGPDE_OPT_GNO_FACTORED = 0)");
      const float* gB = io->phi_params + p.phi.w_off[p.phi.L - 1];
      if (!aligned16(gB)) {
        float* cp = reinterpret_cast<float*>(fws + fp.off_B);
        NGPDE_CUDA_TRY(cudaMemcpyAsync(cp, gB, sizeof(float) * (size_t)R * desc->gno_out, cudaMemcpyDeviceToDevice, st));
        gB = cp;
      }
      if (int rc = gno_gemm(a.gno_S, R, false, gB, desc->gno_out, true, io->mbar,
                            desc->gno_out, g->N, desc->gno_out, R, 1,
                            desc->aggr == NGPDE_AGGR_MEAN ? g->rowptr : nullptr, st))
        return rc;
    }
  }

  // ---- node phase ----
  if (p.has_node) {
    if (int rc = pick_tile([&](int t) { return 4 * fwd_smem(p.node, 0, 0, 0, t).floats; }, &te, &smem)) return rc;
    fs = fwd_smem(p.node, 0, 0, 0, te);
    FwdArgs n{};
    n.tg = TileGraph{nullptr, nullptr, nullptr, nullptr, nullptr, (int)((g->N + te - 1) / te), (int)g->N,
                     (int)std::max<int64_t>(1, g->N / std::max<int64_t>(1, g->G))};
    fill_arrays(g, *desc, p, *io, n.arr, n.ld);
    n.n_segs = p.n_nsegs;
    std::memcpy(n.segs, p.nsegs, sizeof(p.nsegs));
    n.mlp = p.node;
    n.params = io->node_params;
    n.aggr = desc->aggr;
    n.dout = p.dy;
    n.out = io->y;
    n.addend = p.node_addend ? io->mbar : nullptr;
    n.offA = fs.offA; n.offB = fs.offB; n.offW = fs.offW; n.offH = fs.offH;
    ProfScope prof(NGPDE_PROF_FWD_NODE, st);
    if (p.gno_node_gemm) {  // y = act((W x + mbar) + b): one GEMM over the nodes + one elementwise pass
      NGPDE_REQUIRE(aligned16(io->mbar) && aligned16(io->y), "GNOConv node update: mbar and y must be 16-byte aligned");
      if (int rc = layered::run_forward(layered::make_phase(p.node), fp.lygn, fws, fws, layered_gather(g, *desc, p, *io, true), g->N,
                                        io->node_params, false, io->y, nullptr, st, io->mbar))
        return rc;
    } else if (p.nhoist) {
      NGPDE_REQUIRE(fp.node.on, "internal: hoisted node plan without a tensor-core node phase");
      if (int rc = nhoist_forward(g, *desc, p, *io, hbase, nhw, st)) return rc;
      nhoist_args(p, reinterpret_cast<const float*>(hbase + nhw.off_q), &n);
      n.params = reinterpret_cast<const float*>(hbase + nhw.off_fin);
      n.skip_l0 = 1;
      n.tg.n_units = (int)((g->N + TC_TILE - 1) / TC_TILE);
      if (int rc = launch_fwd_tc(true, g->num_sms, fp.node, p.node_in, n.params, n, reinterpret_cast<float*>(fws + fp.node.ws_off), st))
        return rc;
    } else if (fp.node.on) {
      n.tg.n_units = (int)((g->N + TC_TILE - 1) / TC_TILE);
      if (int rc = launch_fwd_tc(true, g->num_sms, fp.node, p.node, io->node_params, n, reinterpret_cast<float*>(fws + fp.node.ws_off), st))
        return rc;
    } else {
      if (int rc = launch_fwd_node(te, n, smem, g->num_sms, st)) return rc;
    }
  }
  return NGPDE_OK;
}

extern "C" int ngpde_conv_backward(ngpde_graph_t g, const ngpde_conv_desc* desc, const ngpde_conv_io* io,
                                   void* workspace, size_t workspace_bytes, void* stream) {
  NGPDE_REQUIRE(g && desc && io, "null argument");
  Plan p;
  if (int rc = make_plan(g, *desc, &p)) return rc;
  if (g->N == 0) return NGPDE_OK;
  if (int rc = check_io(*desc, p, *io, true)) return rc;
  if (p.contract && (desc->aggr == NGPDE_AGGR_MAX || desc->aggr == NGPDE_AGGR_MIN)) {
    set_error("GNOConv backward supports aggr = + and mean only");
    return NGPDE_ERR_UNSUPPORTED;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BwdLayout L;
  if (int rc = bwd_layout(g, *desc, p, &L)) return rc;
  if (workspace_bytes < L.total || workspace == nullptr) {
    set_error("workspace too small: %zu bytes given, %zu needed", workspace_bytes, L.total);
    return NGPDE_ERR_WORKSPACE;
  }
  if (g->N == 0) return NGPDE_OK;
  char* ws = static_cast<char*>(workspace);
  NGPDE_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "workspace must be 16-byte aligned");
  if (p.layered) return layered_backward(g, *desc, p, *io, L, ws, st);
  // hoisted projections: left in io->state by the forward, or recomputed here
  const StatePlan sp = state_plan(p, g->N, g->E);
  const bool kept = io->state != nullptr && sp.bytes > 0;
  NGPDE_REQUIRE(!kept || aligned16(io->state), "io.state must be 16-byte aligned");
  char* hbase = kept ? static_cast<char*>(io->state) : ws;
  const HoistWs& hw = kept ? sp.hoist : L.hoist;
  const NodeHoistWs& nhw = kept ? sp.nhoist : L.nhoist;
  float* wt_phi = reinterpret_cast<float*>(ws + L.off_wt_phi);
  float* wt_node = reinterpret_cast<float*>(ws + L.off_wt_node);
  float* dmbar = p.has_node ? reinterpret_cast<float*>(ws + L.off_dmbar) : nullptr;
  float* dxdirect = p.has_node ? reinterpret_cast<float*>(ws + L.off_dxdirect) : nullptr;
  float* dxdst = (p.edge_dst_side || p.hoist) ? reinterpret_cast<float*>(ws + L.off_dxdst) : nullptr;
  float* desrc = reinterpret_cast<float*>(ws + L.off_desrc);
  float* part_phi = reinterpret_cast<float*>(ws + L.off_part_phi);
  float* part_node = reinterpret_cast<float*>(ws + L.off_part_node);
  float* gS = reinterpret_cast<float*>(ws + L.off_S);
  float* gT = reinterpret_cast<float*>(ws + L.off_T);
  float* gDM = reinterpret_cast<float*>(ws + L.off_DM);
  float* gdB = reinterpret_cast<float*>(ws + L.off_dBpart);
  const int gR = p.gno_Ka * desc->gno_in;
  const float* gB = p.contract == 2 ? io->phi_params + p.phi.w_off[p.phi.L - 1] : nullptr;
  if (p.contract == 2) {
    NGPDE_REQUIRE(aligned16(io->x), "GNOConv (factored evaluation): x must be 16-byte aligned (or set NGPDE_OPT_GNO_FACTORED = 0)");
    if (!aligned16(gB)) {  // the flat parameter segment may start anywhere: the GEMMs read B with 128-bit loads
      float* cp = reinterpret_cast<float*>(ws + L.off_B);
      NGPDE_CUDA_TRY(cudaMemcpyAsync(cp, gB, sizeof(float) * (size_t)gR * desc->gno_out, cudaMemcpyDeviceToDevice, st));
      gB = cp;
    }
  }

  NGPDE_CUDA_TRY(cudaMemsetAsync(part_phi, 0, L.total - L.off_part_phi, st));
  // the transposed weights feed the FFMA kernels' input-gradient GEMMs only (the tensor-core kernels read their own image)
  if (!L.tce.on) transpose_weights_kernel<<<64, 256, 0, st>>>(io->phi_params, wt_phi, p.phi);
  if (p.has_node && !L.tcn.on) transpose_weights_kernel<<<64, 256, 0, st>>>(io->node_params, wt_node, p.node);

  // ---- node phase: dy -> (dx_direct, dmbar, dnode_params) ----
  if (p.has_node) {
    BwdArgs n{};
    const int te = L.tcn.on ? TC_TILE : L.te_n;
    n.tg = TileGraph{nullptr, nullptr, nullptr, nullptr, nullptr, (int)((g->N + te - 1) / te), (int)g->N,
                     (int)std::max<int64_t>(1, g->N / std::max<int64_t>(1, g->G))};
    fill_arrays(g, *desc, p, *io, n.arr, n.ld);
    n.n_segs = p.n_nsegs;
    std::memcpy(n.segs, p.nsegs, sizeof(p.nsegs));
    n.mlp = p.node;
    n.params = io->node_params;
    n.wt = wt_node;
    n.aggr = desc->aggr;
    n.dout = p.dy;
    n.gout_ptr = io->dy;
    n.fwd_out = io->y;
    n.addend = p.node_addend ? io->mbar : nullptr;
    n.dparams_partial = part_node;
    n.part_stride = p.node.n_params;
    n.dx_direct = dxdirect;
    n.dmbar = dmbar;
    n.dx = desc->dx;
    n.need_dz0 = 1;
    n.store_last = L.sn.store_last;
    std::memcpy(n.zoff, L.sn.zoff, sizeof(n.zoff));
    n.offG0 = L.sn.offG0; n.offG1 = L.sn.offG1; n.offW = L.sn.offW;
    if (p.gno_node_gemm) {
      // u = W x + mbar + b recomputed as one GEMM; G = dy * act'(u) is at once the cotangent of mbar; dx_direct = G W', dW = x' G
      ProfScope prof(NGPDE_PROF_BWD_NODE, st);
      NGPDE_REQUIRE(aligned16(io->mbar) && aligned16(io->dy), "GNOConv node update: mbar and dy must be 16-byte aligned");
      const layered::Phase np = layered::make_phase(p.node);
      if (int rc = layered::run_forward(np, L.lygn, ws, ws, layered_gather(g, *desc, p, *io, true), g->N, io->node_params, true, nullptr,
                                        nullptr, st, io->mbar))
        return rc;
      const float *dz0 = nullptr, *gp0 = nullptr;
      if (int rc = layered::run_backward(np, L.lygn, ws, ws, g->N, io->dy, nullptr, nullptr, true, g->num_sms, io->dnode_params, &dz0, st,
                                         &gp0))
        return rc;
      dmbar = const_cast<float*>(gp0);      // [N][dm]
      dxdirect = const_cast<float*>(dz0);   // [N][dx] (dx % 4 == 0: no pad columns)
    } else if (p.nhoist) {
      // first layer of gamma hoisted (Plan::nhoist): inner backward on Q' gives g = d(U + V); the two projections' backward
      // then produce exactly what this phase owes: dx_direct (from U = x Wx + b1) and dmbar (from V = mbar Wm)
      ProfScope prof(NGPDE_PROF_BWD_NODE, st);
      if (!kept) {
        if (int rc = nhoist_forward(g, *desc, p, *io, ws, L.nhoist, st)) return rc;
      }
      float* ndq = reinterpret_cast<float*>(ws + L.off_ndq);
      float* ng = reinterpret_cast<float*>(ws + L.off_ng);
      float* ndfu = reinterpret_cast<float*>(ws + L.off_ndfu);
      float* ndfv = reinterpret_cast<float*>(ws + L.off_ndfv);
      float* ndfin = reinterpret_cast<float*>(ws + L.off_ndfin);
      nhoist_args(p, reinterpret_cast<const float*>(hbase + nhw.off_q), &n);
      n.params = reinterpret_cast<const float*>(hbase + nhw.off_fin);
      n.part_stride = p.node_in.n_params;
      n.skip_w0 = 1;
      n.dx_direct = ndq;   // only columns [0, n1) of a row are written: the cotangent is the same for both halves
      n.dmbar = nullptr;
      n.dx = 2 * p.nh_n1;
      n.tg.n_units = (int)((g->N + TC_TILE - 1) / TC_TILE);
      if (int rc = launch_bwd_tc(true, L.tcn, p.node_in, n, reinterpret_cast<float*>(ws + L.tcn.ws_off), st)) return rc;
      const int Pin = p.node_in.n_params;
      reduce_partials_kernel<<<(Pin + 255) / 256, 256, 0, st>>>(part_node, L.grid_n, Pin, ndfin);
      (void)ng;  // the projections' backward reads the first n1 columns of dQ' in place (strided cotangent)
      const float* fu = reinterpret_cast<const float*>(hbase + nhw.off_fu);
      const float* fv = reinterpret_cast<const float*>(hbase + nhw.off_fv);
      const int ldq = 2 * p.nh_n1;
      if (L.nhgemm && aligned16(io->x) && aligned16(io->mbar)) {
        // GEMM form: dx_direct = g Wx', dmbar = g Wm', dWx = x' g, dWm = mbar' g (split-K, slices in order), db1 = column sums of g
        const int n1 = p.nh_n1;
        float* gpart = reinterpret_cast<float*>(ws + L.off_gpart);
        float* wpart = reinterpret_cast<float*>(ws + L.off_wpart);
        if (int rc = gno_gemm(ndq, ldq, false, fu, n1, false, dxdirect, desc->dx, g->N, desc->dx, n1, 1, nullptr, st)) return rc;
        if (int rc = gno_gemm(ndq, ldq, false, fv, n1, false, dmbar, p.dm, g->N, p.dm, n1, 1, nullptr, st)) return rc;
        if (int rc = node_outer_gemm(io->x, desc->dx, desc->dx, ndq, ldq, n1, g->N, g->num_sms, gpart, ndfu, st)) return rc;
        if (int rc = node_outer_gemm(io->mbar, p.dm, p.dm, ndq, ldq, n1, g->N, g->num_sms, gpart, ndfv, st)) return rc;
        if (int rc = weighted_colsums(ndq, ldq, n1, nullptr, 0, 0, g->N, wpart, ndfu + (size_t)desc->dx * n1, st)) return rc;
      } else {
        if (int rc = node_mlp_backward(g, p.mlp_u, fu, io->x, ndq, dxdirect, ndfu, ws + L.off_nnodews, L.nnodews_bytes, st, nullptr, 0, ldq)) return rc;
        if (int rc = node_mlp_backward(g, p.mlp_v, fv, io->mbar, ndq, dmbar, ndfv, ws + L.off_nnodews, L.nnodews_bytes, st, nullptr, 0, ldq)) return rc;
      }
      nhoist_unfold_kernel<<<32, 256, 0, st>>>(nhoist_map(*desc, p), ndfu, ndfv, ndfin, io->dnode_params);
    } else {
      {
        ProfScope prof(NGPDE_PROF_BWD_NODE, st);
        if (L.tcn.on) {
          n.tg.n_units = (int)((g->N + TC_TILE - 1) / TC_TILE);
          if (int rc = launch_bwd_tc(true, L.tcn, p.node, n, reinterpret_cast<float*>(ws + L.tcn.ws_off), st)) return rc;
        } else {
          if (int rc = launch_bwd_node(te, n, L.smem_n, L.grid_n, st)) return rc;
        }
      }
      const int P = p.node.n_params;
      reduce_partials_kernel<<<(P + 255) / 256, 256, 0, st>>>(part_node, L.grid_n, P, io->dnode_params);
    }
  }

  // ---- edge phase: dmbar -> (dxdst, desrc, dphi_params) ----
  int src_c0_all = 0, src_w_all = desc->dx, dst_c0_all = 0, dst_w_all = desc->dx;
  {
    BwdArgs a{};
    const int te = L.tce.on ? TC_TILE : L.te_e;
    const int ti = tile_index(te);
    a.tg = TileGraph{g->rowptr, g->src, g->dst, g->perm, g->units[ti], g->n_units[ti], (int)g->N,
                     (int)std::max<int64_t>(1, g->E / std::max<int64_t>(1, g->G))};
    fill_arrays(g, *desc, p, *io, a.arr, a.ld);
    a.n_segs = p.n_esegs;
    std::memcpy(a.segs, p.esegs, sizeof(p.esegs));
    a.mlp = p.phi;
    a.params = io->phi_params;
    if (p.hoist) {  // the inner problem on Q (Plan::hoist); its forward part is recomputed here
      if (!kept) {
        if (int rc = hoist_forward(g, *desc, p, *io, ws, L.hoist, st)) return rc;
      }
      a.arr[ARR_X] = reinterpret_cast<const float*>(hbase + hw.off_q);
      a.ld[ARR_X] = 2 * p.h_n1;
      a.n_segs = 2;
      a.segs[0] = p.hsegs[0];
      a.segs[1] = p.hsegs[1];
      a.mlp = p.phi_in;
      a.params = reinterpret_cast<const float*>(hbase + hw.off_fin);
      a.skip_w0 = 1;
    }
    a.wt = wt_phi;
    a.contract = p.contract;
    a.gin = desc->gno_in;
    a.gout = desc->gno_out;
    a.aggr = desc->aggr;
    a.dout = p.dm;
    a.gout_ptr = p.has_node ? dmbar : io->dy;
    a.fwd_out = io->mbar;
    a.dparams_partial = part_phi;
    a.part_stride = L.part_stride;
    a.gno_S = gS; a.gno_T = gT; a.gno_Ka = p.gno_Ka;
    a.offZt = L.se.offZt; a.offTs = L.se.offTs;
    a.debug_skip = g_debug_skip;
    a.dxdst = dxdst;
    a.desrc = desrc;
    a.dx = L.dxe;
    // x columns that receive source-side cotangents (the per-edge spill covers only these on the tensor-core path)
    int src_c0 = 0, src_w = L.dxe, dst_c0 = 0, dst_w = L.dxe;
    if (L.tce.on) {
      int lo = L.dxe, hi = 0;
      for (int i = 0; i < a.n_segs; ++i) {
        const Seg& sg = a.segs[i];
        if (sg.arr != ARR_X || sg.kind == SEG_DST || sg.kind == SEG_EDGE || sg.kind == SEG_GRAPH) continue;
        lo = std::min(lo, sg.col);
        hi = std::max(hi, sg.col + sg.width);
      }
      if (hi > lo) { src_c0 = lo; src_w = hi - lo; }
      lo = L.dxe; hi = 0;
      for (int i = 0; i < a.n_segs; ++i) {
        const Seg& sg = a.segs[i];
        if (sg.arr != ARR_X || !coef_dst_host(sg.kind)) continue;
        lo = std::min(lo, sg.col);
        hi = std::max(hi, sg.col + sg.width);
      }
      if (hi > lo) { dst_c0 = lo; dst_w = hi - lo; }
    }
    a.dst_c0 = dst_c0;
    a.dst_w = dst_w;
    // hoisted input: x' columns [n1, 2 n1) <- input rows [0, n1) through ONE source-kind segment, so an edge's desrc row is its
    // dZ_0 row (needs 16-byte aligned rows: n1 % 4 == 0, which hoisting requires anyway)
    a.direct_src = (p.hoist && L.tce.on && src_w == p.h_n1 && src_c0 == p.h_n1 && (p.h_n1 & 15) == 0) ? 1 : 0;
    a.src_c0 = src_c0;
    a.src_w = src_w;
    src_c0_all = src_c0;
    src_w_all = src_w;
    dst_c0_all = dst_c0;
    dst_w_all = dst_w;
    a.need_dz0 = (p.edge_need_dz0 || p.hoist) ? 1 : 0;
    a.store_last = L.se.store_last;
    a.has_dst_side = (p.edge_dst_side || p.hoist) ? 1 : 0;
    std::memcpy(a.zoff, L.se.zoff, sizeof(a.zoff));
    a.offG0 = L.se.offG0; a.offG1 = L.se.offG1; a.offW = L.se.offW; a.offH = L.se.offH;
    a.offDM = L.se.offDM; a.offP = L.se.offP; a.offDH = L.se.offDH; a.offRed = L.se.offRed;
    if (g->E > 0) {
      ProfScope prof(NGPDE_PROF_BWD_EDGE, st);
      if (p.contract == 2) {
        // DM = dmbar ./ deg;  T = DM B'  (the message cotangent is the same for every in-edge of a node)
        if (int rc = gno_dm_scale(dmbar, g->rowptr, desc->aggr == NGPDE_AGGR_MEAN, g->N, desc->gno_out, gDM, st)) return rc;
        if (int rc = gno_gemm(gDM, desc->gno_out, false, gB, desc->gno_out, false, gT, gR, g->N, gR, desc->gno_out, 1, nullptr, st))
          return rc;
      }
      if (desc->aggr == NGPDE_AGGR_PROD) {
        // the messages once more (forward kernel, nothing aggregated), then their cotangents: dmbar x product of the others
        float* msg = reinterpret_cast<float*>(ws + L.off_msg);
        float* gedge = reinterpret_cast<float*>(ws + L.off_gedge);
        const FwdSmem fs = fwd_smem(p.phi, p.contract, desc->gno_in, desc->gno_out, L.te_f, p.gno_Ka);
        FwdArgs f{};
        const int tf = tile_index(L.te_f);
        f.tg = TileGraph{g->rowptr, g->src, g->dst, g->perm, g->units[tf], g->n_units[tf], (int)g->N, a.tg.gdiv};
        fill_arrays(g, *desc, p, *io, f.arr, f.ld);
        f.n_segs = p.n_esegs;
        std::memcpy(f.segs, p.esegs, sizeof(p.esegs));
        f.mlp = p.phi;
        f.params = io->phi_params;
        f.contract = p.contract;
        f.gin = desc->gno_in;
        f.gout = desc->gno_out;
        f.aggr = desc->aggr;
        f.dout = p.dm;
        f.msg_out = msg;
        f.offA = fs.offA; f.offB = fs.offB; f.offW = fs.offW; f.offH = fs.offH; f.offZt = fs.offZt;
        f.gno_Ka = p.gno_Ka;
        if (int rc = launch_fwd_edge(L.te_f, f, L.smem_f, g->num_sms, st)) return rc;
        const long long tot = (long long)g->N * p.dm;
        prod_cotangent_kernel<<<(unsigned)((tot + 127) / 128), 128, 0, st>>>((int)g->N, p.dm, g->rowptr, msg, a.gout_ptr,
                                                                             reinterpret_cast<float*>(ws + L.off_suf), gedge);
        NGPDE_CUDA_TRY(cudaGetLastError());
        a.gedge = gedge;
      }
      if (L.tce.on) {
        a.tg.unit_ptr = g->units[2];
        a.tg.n_units = g->n_units[2];
        if (int rc = launch_bwd_tc(false, L.tce, a.mlp, a, reinterpret_cast<float*>(ws + L.tce.ws_off), st)) return rc;
      } else if (p.gno_layered) {
        // phi's hidden layers once more as GEMMs (activations kept), the per-destination products (dz, dh, S) with one warp per
        // node, then the hidden layers' backward as GEMMs: their parameter gradients go straight to dphi_params
        const layered::Phase hp = layered::make_phase(p.phi_hidden);
        const int K = p.gno_K;
        layered::Ws w = L.lyg;
        char* kb = ws;
        const float* zbuf = nullptr;
        if (kept) {  // left in io->state by the forward call
          kb = static_cast<char*>(io->state);
          w = sp.lyg;
          w.ga = L.lyg.ga; w.gb = L.lyg.gb; w.part = L.lyg.part; w.cpart = L.lyg.cpart; w.end = L.lyg.end;
          zbuf = reinterpret_cast<const float*>(kb + sp.off_zg);
          gS = reinterpret_cast<float*>(kb + sp.off_Sg);
        } else {     // recomputed: z goes to the cotangent ping-pong buffer run_backward writes last
          if (int rc = layered::run_forward(hp, w, kb, ws, layered_gather(g, *desc, p, *io, false), g->E, io->phi_params, true,
                                            reinterpret_cast<float*>(ws + L.lyg.gb), &zbuf, st))
            return rc;
        }
        float* dzg = reinterpret_cast<float*>(ws + L.off_dzg);
        gnonode::Args na{};
        na.z = zbuf; na.x = io->x; na.ldx = desc->dx; na.src = g->src; na.rowptr = g->rowptr; na.N = (int)g->N; na.Ka = p.gno_Ka;
        na.T = gT; na.S = kept ? nullptr : gS; na.dz = dzg; na.desrc = desrc;  // S kept from the forward: not rebuilt
        if (int rc = gnonode::launch(na, K, true, g->num_sms, st)) return rc;
        if (int rc = layered::run_backward(hp, w, kb, ws, g->E, dzg, nullptr, nullptr, false, g->num_sms, io->dphi_params, nullptr, st))
          return rc;
      } else {
        if (int rc = launch_bwd_edge(te, a, L.smem_e, L.grid_e, st)) return rc;
      }
      if (p.contract == 2) {
        // dB = S' DM over the nodes, split-K slices reduced in fixed order
        if (int rc = gno_gemm(gS, gR, true, gDM, desc->gno_out, true, gdB, desc->gno_out, gR, desc->gno_out, g->N, L.gno_splits, nullptr, st))
          return rc;
        const int PB = gR * desc->gno_out;
        reduce_partials_kernel<<<(PB + 255) / 256, 256, 0, st>>>(gdB, L.gno_splits, PB, io->dphi_params + p.phi.w_off[p.phi.L - 1]);
      }
    } else if (dxdst) {
      NGPDE_CUDA_TRY(cudaMemsetAsync(dxdst, 0, sizeof(float) * g->N * L.dxe, st));
    }
    const int P = L.part_stride;
    float* dphi_target = p.hoist ? reinterpret_cast<float*>(ws + L.off_dfin) : io->dphi_params;
    if (P > 0 && !(p.gno_layered && g->E > 0))
      reduce_partials_kernel<<<(P + 255) / 256, 256, 0, st>>>(part_phi, L.grid_e, P, dphi_target);
    if (p.hoist) {
      // dQ = dxdst + transpose-gather(desrc) -> (dPt, dPs) -> the two projections' backward -> dx, d(Wt, b1, Ws) -> dphi
      ProfScope prof(NGPDE_PROF_BWD_EDGE, st);
      float* dq = reinterpret_cast<float*>(ws + L.off_dq);
      float* dpt = reinterpret_cast<float*>(ws + L.off_dpt);
      float* dps = reinterpret_cast<float*>(ws + L.off_dps);
      float* dxt = reinterpret_cast<float*>(ws + L.off_dxt);
      float* dxs = reinterpret_cast<float*>(ws + L.off_dxs);
      float* dft = reinterpret_cast<float*>(ws + L.off_dft);
      float* dfs = reinterpret_cast<float*>(ws + L.off_dfs);
      const size_t totq = (size_t)g->N * L.dxe;
      if (((L.dxe | src_c0 | src_w | dst_c0 | dst_w) & 3) == 0) {
        dx_combine4_kernel<<<(unsigned)((totq / 4 + 255) / 256), 256, 0, st>>>(
            nullptr, reinterpret_cast<const float4*>(dxdst), g->E > 0 ? reinterpret_cast<const float4*>(desrc) : nullptr, g->tptr, g->tpos,
            (int)g->N, L.dxe / 4, src_c0 / 4, src_w / 4, dst_c0 / 4, dst_w / 4, reinterpret_cast<float4*>(dq));
      } else {
        dx_combine_kernel<<<(unsigned)((totq + 255) / 256), 256, 0, st>>>(nullptr, dxdst, g->E > 0 ? desrc : nullptr, g->tptr, g->tpos,
                                                                           (int)g->N, L.dxe, src_c0, src_w, dst_c0, dst_w, dq);
      }
      (void)dpt; (void)dps;  // dPt / dPs are the two halves of dQ's rows, read in place
      const float* ft = reinterpret_cast<const float*>(hbase + hw.off_ft);
      const float* fs = reinterpret_cast<const float*>(hbase + hw.off_fs);
      const size_t total = (size_t)g->N * desc->dx;
      if (L.hgemm && aligned16(io->x) && (p.ds == 0 || io->snode != nullptr)) {
        // dense GEMMs over the node axis: dx_h = dQ [Wt | Ws]_x' (one GEMM for both projections), d[Wt | Ws]_x = x' dQ (split-K,
        // slices added in order); the static rows and the bias through two-stage weighted column sums
        const int n1 = p.h_n1, hdin = desc->dx + p.ds;
        float* fcat = reinterpret_cast<float*>(ws + L.off_fcat);
        float* dwx = reinterpret_cast<float*>(ws + L.off_dwx);
        float* dsmall = reinterpret_cast<float*>(ws + L.off_dsmall);
        float* gpart = reinterpret_cast<float*>(ws + L.off_gpart);
        float* wpart = reinterpret_cast<float*>(ws + L.off_wpart);
        hoist_cat_kernel<<<(hdin * 2 * n1 + 255) / 256, 256, 0, st>>>(ft, fs, hdin, n1, fcat);
        if (int rc = gno_gemm(dq, L.dxe, false, fcat, 2 * n1, false, dxt, desc->dx, g->N, desc->dx, 2 * n1, 1, nullptr, st)) return rc;
        if (int rc = node_outer_gemm(io->x, desc->dx, desc->dx, dq, L.dxe, 2 * n1, g->N, g->num_sms, gpart, dwx, st)) return rc;
        if (int rc = weighted_colsums(dq, L.dxe, 2 * n1, io->snode, p.ds, p.ds, g->N, wpart, dsmall, st)) return rc;
        hoist_repack_kernel<<<(hdin * n1 + n1 + 255) / 256, 256, 0, st>>>(dwx, dsmall, desc->dx, p.ds, n1, dft, dfs);
        hoist_unfold_kernel<<<32, 256, 0, st>>>(hoist_map(*desc, p), dft, dfs, dphi_target, p.phi.dims[0], io->dphi_params);
        dxs = nullptr;  // both projections' input gradients came out of the one GEMM
      } else {
        if (int rc = node_mlp_backward(g, p.mlp_t, ft, io->x, dq, dxt, dft, ws + L.off_nodews, L.nodews_bytes, st, io->snode, p.ds, L.dxe)) return rc;
        if (int rc = node_mlp_backward(g, p.mlp_s, fs, io->x, dq + p.h_n1, dxs, dfs, ws + L.off_nodews, L.nodews_bytes, st, io->snode, p.ds, L.dxe)) return rc;
        hoist_unfold_kernel<<<32, 256, 0, st>>>(hoist_map(*desc, p), dft, dfs, dphi_target, p.phi.dims[0], io->dphi_params);
      }
      add3_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dxdirect, dxt, dxs, total, io->dx);
      NGPDE_CUDA_TRY(cudaGetLastError());
      return NGPDE_OK;
    }
  }

  // ---- dx = dx_direct + dxdst + transpose-gather(desrc) ----
  {
    const size_t total = (size_t)g->N * desc->dx;
    const bool has_src = (p.edge_need_dz0 || p.contract) && g->E > 0;
    dx_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(dxdirect, dxdst, has_src ? desrc : nullptr,
                                                                        g->tptr, g->tpos, (int)g->N, desc->dx, src_c0_all, src_w_all, dst_c0_all, dst_w_all, io->dx);
  }
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

extern "C" int ngpde_debug_buffer(void* device_int64_x512) {
  tc_set_debug_buffer(static_cast<long long*>(device_int64_x512));
  return NGPDE_OK;
}

extern "C" int ngpde_profile_enable(int32_t on) {
  g_prof_on = on != 0;
  return NGPDE_OK;
}

extern "C" int ngpde_profile_read(double* total_ms, int64_t* launches) {
  NGPDE_REQUIRE(total_ms && launches, "null argument");
  for (int s = 0; s < NGPDE_PROF_SLOTS; ++s) {
    double ms = 0.0;
    for (auto& pr : g_prof[s].ev) {
      NGPDE_CUDA_TRY(cudaEventSynchronize(pr.second));
      float t = 0.f;
      NGPDE_CUDA_TRY(cudaEventElapsedTime(&t, pr.first, pr.second));
      ms += t;
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
    total_ms[s] = ms;
    launches[s] = (int64_t)g_prof[s].ev.size();
    g_prof[s].ev.clear();
  }
  return NGPDE_OK;
}

#define NGPDE_FAMILY_WRAPPER(name, fam)                                                                         \
  extern "C" int ngpde_##name##_forward(ngpde_graph_t g, const ngpde_conv_desc* d, const ngpde_conv_io* io,     \
                                        void* ws, size_t wsb, void* st) {                                       \
    NGPDE_REQUIRE(d && d->family == fam, "descriptor family does not match " #name);                            \
    return ngpde_conv_forward(g, d, io, ws, wsb, st);                                                           \
  }                                                                                                             \
  extern "C" int ngpde_##name##_backward(ngpde_graph_t g, const ngpde_conv_desc* d, const ngpde_conv_io* io,    \
                                         void* ws, size_t wsb, void* st) {                                      \
    NGPDE_REQUIRE(d && d->family == fam, "descriptor family does not match " #name);                            \
    return ngpde_conv_backward(g, d, io, ws, wsb, st);                                                          \
  }

NGPDE_FAMILY_WRAPPER(explicit_edge_conv, NGPDE_EXPLICIT_EDGE_CONV)
NGPDE_FAMILY_WRAPPER(vmh_conv, NGPDE_VMH_CONV)
NGPDE_FAMILY_WRAPPER(mppde_conv, NGPDE_MPPDE_CONV)
NGPDE_FAMILY_WRAPPER(gno_conv, NGPDE_GNO_CONV)
