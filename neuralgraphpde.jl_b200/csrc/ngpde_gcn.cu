// GCNConv (/root/reference/src/layers.jl:200-239), the bare ordered aggregate, and the ODE stage axpy.
//
// The aggregate kernels are the genuinely HBM-bound part of the path: a
// group of d/4 lanes walks one destination row of the merged CSR and accumulates float4 slices of the source
// rows sequentially -- ascending source index, multiply and add rounded separately -- which is the order of the
// dense x SparseMatrixCSC product GNN.jl's CPU path executes.  No atomics; the result is bit-reproducible.
#include <algorithm>

#include "ngpde_conv.cuh"

namespace ngpde {
namespace {

// Per-call edge quantities: merged-entry weights `val` and the normaliser c = 1/sqrt(in-degree).
//   w     : what scales the messages -- the explicit edge_weight, else the graph's own weights when use_edge_weight, else NULL;
//   w_deg : what weights the in-degree -- the explicit edge_weight, else the graph's own stored weights whenever it has
//           any (`degree(g, T; dir=:in, edge_weight)` with edge_weight === nothing resolves to get_edge_weight(g) in
//           GNN.jl's _get_edge_weight, independently of use_edge_weight), else NULL = plain count.
__global__ void gcn_prepare_kernel(int N, int E, int nnz, const int* __restrict__ runptr, const int* __restrict__ order,
                                   const int* __restrict__ rowptr, const int* __restrict__ perm,
                                   const float* __restrict__ w, const float* __restrict__ w_deg, int with_loops,
                                   float* __restrict__ val, float* __restrict__ c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nnz) {
    const int a = runptr[i], b = runptr[i + 1];
    float v;
    if (w == nullptr) {
      v = (float)(b - a);  // duplicates merge into their multiplicity
    } else {
      v = 0.f;
      for (int q = a; q < b; ++q) {
        const int id = order[q];
        v = __fadd_rn(v, id < E ? w[id] : 1.f);  // self-loop weights are padded with ones (layers.jl:215)
      }
    }
    val[i] = v;
  }
  if (i < N) {
    float d;
    if (w_deg == nullptr) {
      d = (float)(rowptr[i + 1] - rowptr[i] + (with_loops ? 1 : 0));  // degree(g, T; dir=:in) of an unweighted graph (layers.jl:224)
    } else {
      d = 0.f;
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) d = __fadd_rn(d, w_deg[perm[k]]);
      if (with_loops) d = __fadd_rn(d, 1.f);
    }
    c[i] = __fdiv_rn(1.f, __fsqrt_rn(d));
  }
}

// out[r][:] = c[r] * sum_{q in ptr[r]..ptr[r+1]} val[ent(q)] * (c[oth(q)] * x[oth(q)][:])
template <int V>
__global__ void __launch_bounds__(256) gcn_aggregate_kernel(int N, int d, const int* __restrict__ ptr,
                                                             const int* __restrict__ ent, const int* __restrict__ other,
                                                             const float* __restrict__ val, const float* __restrict__ c,
                                                             const float* __restrict__ x, float* __restrict__ out) {
  const int lpr = d / V;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long r64 = idx / lpr;
  if (r64 >= N) return;
  const int r = (int)r64, lane = (int)(idx - r64 * lpr);
  float acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;
  const int q0 = ptr[r], q1 = ptr[r + 1];
  // Four entries per trip: their index / weight / normaliser loads and then their source rows are all in flight before the
  // first add (the kernel is bound by memory latency, not by issue), and the adds still run in ascending entry order.
  constexpr int U = 4;
  for (int q = q0; q < q1; q += U) {
    int sidx[U];
    float w[U], cs[U];
    float xv[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool on = q + u < q1;
      const int j = on ? (ent ? __ldg(ent + q + u) : q + u) : 0;
      sidx[u] = on ? __ldg(other + j) : -1;
      w[u] = on ? __ldg(val + j) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (sidx[u] >= 0) {
        cs[u] = __ldg(c + sidx[u]);
        if (V == 4) {
          *reinterpret_cast<float4*>(xv[u]) = __ldg(reinterpret_cast<const float4*>(x + (size_t)sidx[u] * d + lane * 4));
        } else {
          xv[u][0] = __ldg(x + (size_t)sidx[u] * d + lane);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (sidx[u] >= 0) {
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] = __fadd_rn(acc[v], __fmul_rn(__fmul_rn(xv[u][v], cs[u]), w[u]));
      }
    }
  }
  const float cr = c[r];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = __fmul_rn(acc[v], cr);
  if (V == 4) {
    *reinterpret_cast<float4*>(out + (size_t)r * d + lane * 4) = *reinterpret_cast<float4*>(acc);
  } else {
    out[(size_t)r * d + lane] = acc[0];
  }
}

// The same sum for rows that are a whole number of float4, with the row width a compile-time constant: no 64-bit division to
// find (row, lane), one IMAD.WIDE per source row, four entries in flight per trip and a branch-free body for full groups.
// (The generic kernel above spends most of its issue slots on index arithmetic: 75 % issue utilisation at 19 % of the HBM
// peak, profiles/r02k_ncu_gcn_summary.md.)  Entry order and roundings are unchanged: ((x * c_s) * w) added in ascending order.
template <int LPR, bool HAS_ENT>
__global__ void __launch_bounds__(256) gcn_aggregate_v4_kernel(int N, const int* __restrict__ ptr, const int* __restrict__ ent,
                                                                const int* __restrict__ other, const float* __restrict__ val,
                                                                const float* __restrict__ c, const float4* __restrict__ x,
                                                                float4* __restrict__ out) {
  const unsigned gid = blockIdx.x * 256u + threadIdx.x;
  const unsigned r = gid / LPR, lane = gid % LPR;
  if (r >= (unsigned)N) return;
  const float4* xl = x + lane;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  auto add = [&](const float4 v, const float cs, const float w) {
    float4 t = make_float4(__fmul_rn(v.x, cs), __fmul_rn(v.y, cs), __fmul_rn(v.z, cs), __fmul_rn(v.w, cs));
    if (w != 1.f) {  // an unweighted, duplicate-free entry has val == 1: multiplying by it is exact, so it is skipped
      t.x = __fmul_rn(t.x, w); t.y = __fmul_rn(t.y, w); t.z = __fmul_rn(t.z, w); t.w = __fmul_rn(t.w, w);
    }
    acc.x = __fadd_rn(acc.x, t.x);
    acc.y = __fadd_rn(acc.y, t.y);
    acc.z = __fadd_rn(acc.z, t.z);
    acc.w = __fadd_rn(acc.w, t.w);
  };
  int q = __ldg(ptr + r);
  const int q1 = __ldg(ptr + r + 1);
  for (; q + 4 <= q1; q += 4) {
    int j[4], s[4];
    float w[4], cs[4];
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) j[u] = HAS_ENT ? __ldg(ent + q + u) : q + u;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      s[u] = __ldg(other + j[u]);
      w[u] = __ldg(val + j[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      cs[u] = __ldg(c + s[u]);
      v[u] = __ldg(xl + (size_t)s[u] * LPR);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) add(v[u], cs[u], w[u]);
  }
  for (; q < q1; ++q) {
    const int j = HAS_ENT ? __ldg(ent + q) : q;
    const int s = __ldg(other + j);
    add(__ldg(xl + (size_t)s * LPR), __ldg(c + s), __ldg(val + j));
  }
  const float cr = __ldg(c + r);
  acc.x = __fmul_rn(acc.x, cr);
  acc.y = __fmul_rn(acc.y, cr);
  acc.z = __fmul_rn(acc.z, cr);
  acc.w = __fmul_rn(acc.w, cr);
  out[(size_t)r * LPR + lane] = acc;
}

// Same sum once more (developer variant, NOT the default: see gcn_variant), for Blackwell's packed FP32 pipe: `mul.rn.f32x2` / `add.rn.f32x2` round each half exactly like the scalar
// instructions (so entry order and roundings are STILL ((x * c_s) * w) added in ascending order) at half the issue slots,
// and U entries per trip are in flight, predicated instead of split into a full-group loop and a tail.
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

template <int LPR, bool HAS_ENT, int U>
__global__ void __launch_bounds__(256) gcn_aggregate_v5_kernel(int N, const int* __restrict__ ptr, const int* __restrict__ ent,
                                                                const int* __restrict__ other, const float* __restrict__ val,
                                                                const float* __restrict__ c, const float4* __restrict__ x,
                                                                float4* __restrict__ out) {
  const unsigned gid = blockIdx.x * 256u + threadIdx.x;
  const unsigned r = gid / LPR, lane = gid % LPR;
  if (r >= (unsigned)N) return;
  const float4* xl = x + lane;
  unsigned long long a01 = 0ull, a23 = 0ull;  // (+0, +0)
  int q = __ldg(ptr + r);
  const int q1 = __ldg(ptr + r + 1);
  const float cr = __ldg(c + r);
  for (; q < q1; q += U) {
    int s[U];
    float w[U], cs[U];
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool on = q + u < q1;
      const int j = on ? (HAS_ENT ? __ldg(ent + q + u) : q + u) : 0;
      s[u] = on ? __ldg(other + j) : -1;
      w[u] = on ? __ldg(val + j) : 1.f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (s[u] >= 0) {
        cs[u] = __ldg(c + s[u]);
        v[u] = __ldg(xl + (size_t)s[u] * LPR);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (s[u] >= 0) {
        const unsigned long long c2 = f2_pack(cs[u], cs[u]);
        unsigned long long t01 = f2_mul(f2_pack(v[u].x, v[u].y), c2), t23 = f2_mul(f2_pack(v[u].z, v[u].w), c2);
        if (w[u] != 1.f) {  // an unweighted, duplicate-free entry has val == 1: multiplying by it is exact, so it is skipped
          const unsigned long long w2 = f2_pack(w[u], w[u]);
          t01 = f2_mul(t01, w2);
          t23 = f2_mul(t23, w2);
        }
        a01 = f2_add(a01, t01);
        a23 = f2_add(a23, t23);
      }
    }
  }
  const unsigned long long cr2 = f2_pack(cr, cr);
  a01 = f2_mul(a01, cr2);
  a23 = f2_mul(a23, cr2);
  float4 o;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(a01));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o.z), "=f"(o.w) : "l"(a23));
  out[(size_t)r * LPR + lane] = o;
}

// developer switch NGPDE_GCN_V: 0 (default) scalar kernel with 4 entries in flight; 1 / 2 the packed kernel with 4 / 8 in flight.
// Measured on the C5 graph (2M nodes, d = 64, profiles/r02p_gcn_variants.csv): 420 / 612 / 870 us -- the packed kernels need 48 /
// 80 registers against 32, and this latency-bound gather wants resident warps more than it wants issue slots.
int gcn_variant() {
  const char* e = getenv("NGPDE_GCN_V");
  return e ? atoi(e) : 0;
}

template <int LPR>
int launch_gcn_v4(int N, const int* ptr, const int* ent, const int* other, const float* val, const float* c, const float* x,
                  float* out, cudaStream_t st) {
  const unsigned blocks = (unsigned)(((long long)N * LPR + 255) / 256);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float4* o4 = reinterpret_cast<float4*>(out);
  const int variant = gcn_variant();
  if (variant == 2) {
    if (ent) gcn_aggregate_v5_kernel<LPR, true, 8><<<blocks, 256, 0, st>>>(N, ptr, ent, other, val, c, x4, o4);
    else gcn_aggregate_v5_kernel<LPR, false, 8><<<blocks, 256, 0, st>>>(N, ptr, ent, other, val, c, x4, o4);
  } else if (variant == 1) {
    if (ent) gcn_aggregate_v5_kernel<LPR, true, 4><<<blocks, 256, 0, st>>>(N, ptr, ent, other, val, c, x4, o4);
    else gcn_aggregate_v5_kernel<LPR, false, 4><<<blocks, 256, 0, st>>>(N, ptr, ent, other, val, c, x4, o4);
  } else if (ent) gcn_aggregate_v4_kernel<LPR, true><<<blocks, 256, 0, st>>>(N, ptr, ent, other, val, c, x4, o4);
  else gcn_aggregate_v4_kernel<LPR, false><<<blocks, 256, 0, st>>>(N, ptr, ent, other, val, c, x4, o4);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

int launch_gcn_aggregate(int N, int d, const int* ptr, const int* ent, const int* other, const float* val,
                         const float* c, const float* x, float* out, cudaStream_t st) {
  if (N == 0) return NGPDE_OK;
  const bool v4 = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  if (v4 && (long long)N * (d / 4) < (1ll << 31)) {
    switch (d / 4) {
      case 1: return launch_gcn_v4<1>(N, ptr, ent, other, val, c, x, out, st);
      case 2: return launch_gcn_v4<2>(N, ptr, ent, other, val, c, x, out, st);
      case 4: return launch_gcn_v4<4>(N, ptr, ent, other, val, c, x, out, st);
      case 8: return launch_gcn_v4<8>(N, ptr, ent, other, val, c, x, out, st);
      case 16: return launch_gcn_v4<16>(N, ptr, ent, other, val, c, x, out, st);
      case 32: return launch_gcn_v4<32>(N, ptr, ent, other, val, c, x, out, st);
      default: break;
    }
  }
  const long long threads = (long long)N * (v4 ? d / 4 : d);
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  if (v4) gcn_aggregate_kernel<4><<<blocks, 256, 0, st>>>(N, d, ptr, ent, other, val, c, x, out);
  else gcn_aggregate_kernel<1><<<blocks, 256, 0, st>>>(N, d, ptr, ent, other, val, c, x, out);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

// y = act(a + b) over [N][D]  (the out < in branch applies bias and activation after the aggregate)
__global__ void bias_act_kernel(const float* __restrict__ a, const float* __restrict__ bias, int act, size_t total, int D,
                                float* __restrict__ y) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float b = bias ? bias[i % D] : 0.f;
  y[i] = act_fwd(act, a[i] + b);
}

// dP = dy * act'(a + b)
__global__ void bias_act_grad_kernel(const float* __restrict__ a, const float* __restrict__ bias, int act, size_t total,
                                     int D, const float* __restrict__ dy, float* __restrict__ dp) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float b = bias ? bias[i % D] : 0.f;
  dp[i] = dy[i] * act_grad_pre(act, a[i] + b);
}

// deterministic column sums of [N][D]: stage 1 -> partial[nblk][D], stage 2 -> out[D]
__global__ void colsum_stage1_kernel(const float* __restrict__ a, int N, int D, int rows_per_block, float* __restrict__ partial) {
  const int r0 = blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += a[(size_t)r * D + c];
    partial[(size_t)blockIdx.x * D + c] = s;
  }
}
__global__ void colsum_stage2_kernel(const float* __restrict__ partial, int nblk, int D, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partial[(size_t)b * D + c];
  out[c] = s;
}

// bare propagate(copy_xj | e_mul_xj, g, aggr): thread per (row, channel), sequential over the CSR row
__global__ void aggregate_kernel(int N, int d, int aggr, const int* __restrict__ rowptr, const int* __restrict__ src,
                                 const int* __restrict__ perm, const float* __restrict__ x, const float* __restrict__ w,
                                 float* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * d) return;
  const int r = (int)(idx / d), c = (int)(idx - (long long)r * d);
  const int q0 = rowptr[r], q1 = rowptr[r + 1];
  float acc = aggr == NGPDE_AGGR_MAX ? -INFINITY : (aggr == NGPDE_AGGR_MIN ? INFINITY : (aggr == NGPDE_AGGR_PROD ? 1.f : 0.f));
  for (int q = q0; q < q1; ++q) {
    float v = x[(size_t)src[q] * d + c];
    if (w) v = __fmul_rn(w[perm[q]], v);
    if (aggr == NGPDE_AGGR_MAX) acc = fmaxf(acc, v);
    else if (aggr == NGPDE_AGGR_MIN) acc = fminf(acc, v);
    else if (aggr == NGPDE_AGGR_PROD) acc = __fmul_rn(acc, v);
    else acc = __fadd_rn(acc, v);
  }
  if (aggr == NGPDE_AGGR_MEAN && q1 > q0) acc = __fdiv_rn(acc, (float)(q1 - q0));
  out[idx] = acc;
}

__global__ void axpy_stages_kernel(float* __restrict__ out, const float* __restrict__ u, const float* k0, const float* k1,
                                   const float* k2, const float* k3, const float* k4, const float* k5, const float* k6,
                                   const float* k7, float c0, float c1, float c2, float c3, float c4, float c5, float c6,
                                   float c7, int nk, long long n) {
  const float* ks[8] = {k0, k1, k2, k3, k4, k5, k6, k7};
  const float cs[8] = {c0, c1, c2, c3, c4, c5, c6, c7};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = u ? u[i] : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < nk) v = fmaf(cs[j], ks[j][i], v);
    out[i] = v;
  }
}

size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct GcnWs {
  size_t off_val, off_c, off_agg, off_lin, off_tmp, off_tmp2, off_tmp3, off_cs, off_wblk, wblk_bytes, off_mlp, total;
  int dmin;
  int cs_blocks, cs_rows;
};

int gcn_mlp(const ngpde_gcn_desc& d, bool first, MlpDev* m) {
  ngpde_mlp h{};
  h.n_layers = 1;
  h.dims[0] = d.in_chs;
  h.dims[1] = d.out_chs;
  // out >= in: the Dense carries bias and activation; out < in: W is applied first, bare.
  h.act[0] = first ? NGPDE_ACT_IDENTITY : d.act;
  h.has_bias[0] = first ? 0 : d.has_bias;
  return make_mlp_dev(h, m, "GCNConv weight");
}

// activation the Dense backward of the out >= in branch runs with: identity when act' is a function of the output (the
// cotangent is pre-multiplied by act'(y) in one elementwise pass), else the activation itself (swish / gelu)
int gcn_bwd_act(int act) { return act_grad_from_y(act) ? NGPDE_ACT_IDENTITY : act; }

__global__ void act_grad_y_kernel(const float* __restrict__ y, const float* __restrict__ dy, int act, size_t total, float* __restrict__ dp) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) dp[i] = dy[i] * act_grad_y(act, y[i]);
}

int gcn_ws(const ngpde_graph* g, const ngpde_gcn_desc& d, bool backward, GcnWs* w) {
  const GcnLayout& L = g->gcn[d.add_self_loops ? 1 : 0];
  const size_t N = (size_t)g->N;
  w->dmin = std::min(d.in_chs, d.out_chs);
  size_t off = 0;
  w->off_val = off; off = align256(off + sizeof(float) * std::max(L.nnz, 1));
  w->off_c = off;   off = align256(off + sizeof(float) * std::max<size_t>(N, 1));
  w->off_agg = off; off = align256(off + sizeof(float) * N * w->dmin);
  w->off_lin = off; off = align256(off + sizeof(float) * N * w->dmin);
  w->off_tmp = off; off = align256(off + (backward ? sizeof(float) * N * std::max(d.in_chs, d.out_chs) : 0));
  w->off_tmp2 = off; off = align256(off + (backward ? sizeof(float) * N * w->dmin : 0));
  w->cs_rows = 256;
  w->cs_blocks = (int)((N + w->cs_rows - 1) / w->cs_rows);
  w->off_cs = off;  off = align256(off + (backward ? sizeof(float) * (size_t)std::max(w->cs_blocks, 1) * d.out_chs : 0));
  {
    const bool first = d.out_chs < d.in_chs;
    MlpDev m;
    if (int rc = gcn_mlp(d, first, &m)) return rc;
    w->off_wblk = off;  // prepared weight block of the tensor-core Dense (0 bytes: FFMA engine)
    w->wblk_bytes = node_mlp_forward_ws(m);
    off = align256(off + w->wblk_bytes);
    // out >= in, backward: dP = dy * act'(y) so that the Dense backward sees an identity activation (tensor-core eligible)
    w->off_tmp3 = off;
    off = align256(off + ((backward && !first) ? sizeof(float) * N * d.out_chs : 0));
    w->off_mlp = off;
    if (backward) {
      m.act[0] = first ? m.act[0] : gcn_bwd_act(d.act);
      off = align256(off + node_mlp_backward_ws(g, m));
    }
  }
  w->total = off;
  return NGPDE_OK;
}

int gcn_check(const ngpde_graph* g, const ngpde_gcn_desc* d) {
  NGPDE_REQUIRE(g && d, "null argument");
  NGPDE_REQUIRE(d->in_chs > 0 && d->out_chs > 0, "GCNConv channel counts must be positive");
  NGPDE_REQUIRE(d->act >= NGPDE_ACT_IDENTITY && d->act <= NGPDE_ACT_LEAKYRELU, "unknown activation %d", d->act);
  return NGPDE_OK;
}

}  // namespace
}  // namespace ngpde

using namespace ngpde;

extern "C" size_t ngpde_gcn_workspace_bytes(ngpde_graph_t g, const ngpde_gcn_desc* desc, int32_t backward) {
  if (gcn_check(g, desc)) return 0;
  if (build_gcn_layout(g, desc->add_self_loops, 0)) return 0;
  GcnWs w;
  if (gcn_ws(g, *desc, backward != 0, &w)) return 0;
  return w.total + 256;
}

extern "C" int ngpde_gcn_conv_forward(ngpde_graph_t g, const ngpde_gcn_desc* desc, const float* x, const float* weight,
                                      const float* bias, const float* edge_weight, const float* graph_weight, float* y,
                                      void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = gcn_check(g, desc)) return rc;
  NGPDE_REQUIRE(x && weight && y, "null tensor argument");
  NGPDE_REQUIRE(!desc->has_bias || bias, "bias is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = build_gcn_layout(g, desc->add_self_loops, st)) return rc;
  const GcnLayout& L = g->gcn[desc->add_self_loops ? 1 : 0];
  GcnWs w;
  if (int rc = gcn_ws(g, *desc, false, &w)) return rc;
  if (workspace == nullptr || workspace_bytes < w.total) {
    set_error("GCNConv workspace too small: %zu < %zu", workspace_bytes, w.total);
    return NGPDE_ERR_WORKSPACE;
  }
  if (g->N == 0) return NGPDE_OK;
  char* ws = static_cast<char*>(workspace);
  float* val = reinterpret_cast<float*>(ws + w.off_val);
  float* c = reinterpret_cast<float*>(ws + w.off_c);
  float* agg = reinterpret_cast<float*>(ws + w.off_agg);
  float* lin = reinterpret_cast<float*>(ws + w.off_lin);
  const int N = (int)g->N, E = (int)g->E;
  const float* wmsg = edge_weight ? edge_weight : (desc->use_edge_weight ? graph_weight : nullptr);
  const float* wdeg = edge_weight ? edge_weight : graph_weight;
  gcn_prepare_kernel<<<(std::max(N, L.nnz) + 255) / 256, 256, 0, st>>>(N, E, L.nnz, L.runptr, L.order, g->rowptr, g->perm,
                                                                     wmsg, wdeg, desc->add_self_loops, val, c);
  const bool first = desc->out_chs < desc->in_chs;
  MlpDev m;
  if (int rc = gcn_mlp(*desc, first, &m)) return rc;
  if (first) {
    if (int rc = node_mlp_forward(g, m, weight, x, lin, st, ws + w.off_wblk, w.wblk_bytes)) return rc;  // layers.jl:220-223
    if (int rc = launch_gcn_aggregate(N, desc->out_chs, L.colptr, nullptr, L.rowval, val, c, lin, agg, st)) return rc;
    const size_t total = (size_t)N * desc->out_chs;
    bias_act_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(agg, desc->has_bias ? bias : nullptr, desc->act, total,
                                                                     desc->out_chs, y);
  } else {
    if (int rc = launch_gcn_aggregate(N, desc->in_chs, L.colptr, nullptr, L.rowval, val, c, x, agg, st)) return rc;
    // weight and bias are adjacent in the flat parameter vector; pass them as one segment when they are
    NGPDE_REQUIRE(!desc->has_bias || bias == weight + (size_t)desc->in_chs * desc->out_chs,
                  "GCNConv expects bias to follow weight in the flat parameter vector");
    if (int rc = node_mlp_forward(g, m, weight, agg, y, st, ws + w.off_wblk, w.wblk_bytes)) return rc;  // layers.jl:235-238
  }
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

extern "C" int ngpde_gcn_conv_backward(ngpde_graph_t g, const ngpde_gcn_desc* desc, const float* x, const float* weight,
                                       const float* bias, const float* edge_weight, const float* graph_weight,
                                       const float* y, const float* dy, float* dx, float* dweight, float* dbias,
                                       void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = gcn_check(g, desc)) return rc;
  NGPDE_REQUIRE(x && weight && dy && dx && dweight, "null tensor argument");
  NGPDE_REQUIRE(!desc->has_bias || (bias && dbias), "bias/dbias is NULL");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = build_gcn_layout(g, desc->add_self_loops, st)) return rc;
  const GcnLayout& L = g->gcn[desc->add_self_loops ? 1 : 0];
  GcnWs w;
  if (int rc = gcn_ws(g, *desc, true, &w)) return rc;
  if (workspace == nullptr || workspace_bytes < w.total) {
    set_error("GCNConv workspace too small: %zu < %zu", workspace_bytes, w.total);
    return NGPDE_ERR_WORKSPACE;
  }
  if (g->N == 0) return NGPDE_OK;
  char* ws = static_cast<char*>(workspace);
  float* val = reinterpret_cast<float*>(ws + w.off_val);
  float* c = reinterpret_cast<float*>(ws + w.off_c);
  float* agg = reinterpret_cast<float*>(ws + w.off_agg);
  float* lin = reinterpret_cast<float*>(ws + w.off_lin);
  float* tmp = reinterpret_cast<float*>(ws + w.off_tmp);
  float* tmp2 = reinterpret_cast<float*>(ws + w.off_tmp2);
  float* cs = reinterpret_cast<float*>(ws + w.off_cs);
  void* mlp_ws = ws + w.off_mlp;
  const size_t mlp_ws_bytes = w.total - w.off_mlp;
  const int N = (int)g->N, E = (int)g->E;
  const float* wmsg = edge_weight ? edge_weight : (desc->use_edge_weight ? graph_weight : nullptr);
  const float* wdeg = edge_weight ? edge_weight : graph_weight;
  // recompute the forward intermediates (val, c, aggregate): cheaper than keeping them alive between calls
  gcn_prepare_kernel<<<(std::max(N, L.nnz) + 255) / 256, 256, 0, st>>>(N, E, L.nnz, L.runptr, L.order, g->rowptr, g->perm,
                                                                     wmsg, wdeg, desc->add_self_loops, val, c);
  const bool first = desc->out_chs < desc->in_chs;
  MlpDev m;
  if (int rc = gcn_mlp(*desc, first, &m)) return rc;
  if (first) {
    // y = act(A(Wx) + b):  dP = dy*act'(.), db = colsum(dP), du = A^T dP, (dW, dx) from the bare Dense
    if (int rc = node_mlp_forward(g, m, weight, x, lin, st, ws + w.off_wblk, w.wblk_bytes)) return rc;
    if (int rc = launch_gcn_aggregate(N, desc->out_chs, L.colptr, nullptr, L.rowval, val, c, lin, agg, st)) return rc;
    const size_t total = (size_t)N * desc->out_chs;
    bias_act_grad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(agg, desc->has_bias ? bias : nullptr, desc->act,
                                                                          total, desc->out_chs, dy, tmp);
    if (desc->has_bias) {
      colsum_stage1_kernel<<<std::max(w.cs_blocks, 1), 128, 0, st>>>(tmp, N, desc->out_chs, w.cs_rows, cs);
      colsum_stage2_kernel<<<(desc->out_chs + 127) / 128, 128, 0, st>>>(cs, w.cs_blocks, desc->out_chs, dbias);
    }
    if (int rc = launch_gcn_aggregate(N, desc->out_chs, L.tptr, L.tpos, L.colidx, val, c, tmp, tmp2, st)) return rc;
    if (int rc = node_mlp_backward(g, m, weight, x, tmp2, dx, dweight, mlp_ws, mlp_ws_bytes, st)) return rc;
  } else {
    // y = act(W A(x) + b): Dense backward gives (dW, db, dA); dx = A^T dA
    NGPDE_REQUIRE(!desc->has_bias || (bias == weight + (size_t)desc->in_chs * desc->out_chs &&
                                     dbias == dweight + (size_t)desc->in_chs * desc->out_chs),
                  "GCNConv expects bias to follow weight in the flat parameter vector (and dbias to follow dweight)");
    if (int rc = launch_gcn_aggregate(N, desc->in_chs, L.colptr, nullptr, L.rowval, val, c, x, agg, st)) return rc;
    const float* dp = dy;
    const float* yact = nullptr;
    if (gcn_bwd_act(desc->act) != desc->act) {
      NGPDE_REQUIRE(y != nullptr, "GCNConv backward needs the forward output y");
      m.act[0] = NGPDE_ACT_IDENTITY;
      if (node_mlp_backward_fuses_act(g, m)) {
        yact = y;  // the tensor-core Dense backward multiplies the cotangent by act'(y) as it loads it
      } else {
        const size_t total = (size_t)N * desc->out_chs;
        float* tmp3 = reinterpret_cast<float*>(ws + w.off_tmp3);
        act_grad_y_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(y, dy, desc->act, total, tmp3);
        dp = tmp3;
      }
    }
    if (int rc = node_mlp_backward(g, m, weight, agg, dp, tmp, dweight, mlp_ws, mlp_ws_bytes, st, nullptr, 0, 0, yact, desc->act)) return rc;
    if (int rc = launch_gcn_aggregate(N, desc->in_chs, L.tptr, L.tpos, L.colidx, val, c, tmp, dx, st)) return rc;
  }
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

extern "C" int ngpde_aggregate(ngpde_graph_t g, int32_t aggr, const float* x, int32_t d, const float* w, float* out,
                               void* stream) {
  NGPDE_REQUIRE(g && x && out && d > 0, "bad argument");
  NGPDE_REQUIRE(aggr >= NGPDE_AGGR_SUM && aggr <= NGPDE_AGGR_PROD, "unknown aggregation %d", aggr);
  if (g->N == 0) return NGPDE_OK;
  const long long total = (long long)g->N * d;
  aggregate_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (int)g->N, d, aggr, g->rowptr, g->src, g->perm, x, w, out);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

extern "C" int ngpde_axpy_stages(float* out, const float* u, const float* const* k, const float* coef, int32_t nk,
                                 int64_t n, void* stream) {
  NGPDE_REQUIRE(out && nk >= 0 && nk <= 8 && n >= 0, "bad argument");
  NGPDE_REQUIRE(nk == 0 || (k && coef), "null stage arrays");
  if (n == 0) return NGPDE_OK;
  const float* ks[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  float cs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < nk; ++j) { ks[j] = k[j]; cs[j] = coef[j]; }
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
  axpy_stages_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      out, u, ks[0], ks[1], ks[2], ks[3], ks[4], ks[5], ks[6], ks[7], cs[0], cs[1], cs[2], cs[3], cs[4], cs[5], cs[6],
      cs[7], nk, n);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}
