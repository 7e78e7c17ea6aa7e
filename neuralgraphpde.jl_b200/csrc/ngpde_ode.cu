// Persistent fixed-step Runge-Kutta integrator (and its discrete adjoint) for  du/dt = ExplicitEdgeConv(u)  on a graph small
// enough that one right-hand side is launch latency, not work -- SURVEY.md section 8f-1, BASELINE config C1 (1,024 nodes,
// 3,968 edges, phi 4 => 16 => 16 => 1): the reference's loop `solve(prob, Tsit5(); adaptive = false, dt)` around
// `dudt(u, p, t) = model(u, p, st)[1]` (docs/src/tutorials/graph_node.md:53-66) makes ~15 library launches per RHS and
// 121 RHS per trajectory.
//
// Here ONE kernel integrates all steps.  A thread-block cluster of ODE_CTAS CTAs owns the graph: CTA c owns a contiguous
// node range and the in-edges of those nodes (CSR order).  Per right-hand side: every thread evaluates phi for its edges in
// registers (weights broadcast from shared memory), the messages of a destination are added in ascending CSR position
// (the order of NNlib's scatter: bit-identical to the layer kernels' aggregation), the owning thread immediately forms
// the next stage's input  u + dt sum_j a_sj k_j  and ONE cluster barrier makes it visible to the neighbours' gathers.
// The adjoint kernel walks the steps backwards from the saved stage inputs: per stage it recomputes phi on 128-edge tiles
// kept in shared memory, back-propagates, adds the parameter gradient with a fixed thread <-> parameter ownership and a
// fixed edge order (deterministic, no atomics), spills the source-side input cotangents per edge and gathers them over the
// transpose after one cluster barrier.
#include <cooperative_groups.h>

#include <algorithm>
#include <cstring>

#include "ngpde_conv.cuh"

namespace cg = cooperative_groups;

namespace ngpde {
namespace {

constexpr int ODE_CTAS = 8;       // portable cluster size
constexpr int ODE_THREADS = 512;
constexpr int ODE_MAXW = 32;      // widest phi layer the kernels are instantiated for
constexpr int ODE_MAXIN = 16;     // widest phi input
constexpr int ODE_MAXS = 8;       // Runge-Kutta stages

struct OdeArgs {
  const int* rowptr;
  const int* src;
  const int* dst;
  const int* tptr;
  const int* tpos;
  int N, E;
  int dx, dhs, dpos, din;     // state width, static columns riding with h, position columns, phi input width
  int aggr;
  MlpDev mlp;
  const float* params;
  const float* snode;         // [N][dhs + dpos]
  int S;
  float a[ODE_MAXS][ODE_MAXS];
  float b[ODE_MAXS];
  float dt;
  int n_steps;
  float* u;                   // [N][dx] state, in/out (forward); unused by the adjoint
  float* traj;                // [n_steps][S][N][dx] stage inputs
  float* kbuf;                // [S][N][dx] stage derivatives (forward) / stage input cotangents ubar (adjoint)
  float* lam;                 // adjoint: [N][dx] in/out
  float* desrc;               // adjoint: [2][E][dx] per-edge source-side cotangents (double buffered across stages)
  float* dpart;               // adjoint: [ODE_CTAS][n_params]
  float* dparams;             // adjoint: [n_params]
  int max_edges;              // largest edge count of any CTA (sizes the per-CTA message buffer)
};

// tanh(x) = 1 - 2 / (2^(2x log2 e) + 1) on the SFU exponential and reciprocal: absolute error <= 1.5e-7 (the tensor-core
// kernels' tc_tanh); every other activation takes the accurate library form
__device__ __forceinline__ float ode_act(int a, float x) {
  if (a == NGPDE_ACT_TANH) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    return fmaf(-2.f, r, 1.f);
  }
  if (a == NGPDE_ACT_IDENTITY) return x;
  if (a == NGPDE_ACT_RELU) return fmaxf(x, 0.f);
  return act_fwd(a, x);
}

// Parameters in CONSTANT memory, padded: layer l is a [W][W] block (row k = input k, zero beyond the layer's real shape)
// followed by W biases.  Every thread needs every weight: from shared memory that is one broadcast load per 4 FMAs and the
// load pipe bounds the kernel (measured: 16k cycles per right-hand side at C1); from the constant bank the weight is an
// immediate operand of the FFMA itself.  Zero padding makes every loop bound a compile-time constant: padded outputs are
// act(0), but they only ever meet zero weights downstream, and padded cotangents are exactly 0.
// (One module-wide buffer: launches of these kernels with DIFFERENT parameters must be ordered on one stream.)
constexpr int ODE_MAXL = 4;
__constant__ float c_ode_w[ODE_MAXL * (ODE_MAXW * ODE_MAXW + ODE_MAXW)];

__global__ void ode_pad_params_kernel(MlpDev m, const float* __restrict__ params, int W, float* __restrict__ wp) {
  const int LS = W * W + W;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m.L * LS; i += gridDim.x * blockDim.x) {
    const int l = i / LS, r = i - l * LS;
    const int K = m.dims[l], N = m.dims[l + 1];
    float v = 0.f;
    if (r < W * W) {
      const int k = r / W, n = r - k * W;
      if (k < K && n < N) v = params[m.w_off[l] + k * N + n];
    } else {
      const int n = r - W * W;
      if (n < N && m.b_off[l] >= 0) v = params[m.b_off[l] + n];
    }
    wp[i] = v;
  }
}

// phi on one edge, all in registers; the layer loop is unrolled so that every weight is a compile-time constant-bank
// address.  KEEP (adjoint): the inputs of layers 1.. are also written to the tile (row-major [layer][TE][W + 1], `zs`
// points at this edge's row of layer 0, `lstride` floats between layers).
template <int W, bool KEEP>
__device__ __forceinline__ void ode_mlp(const MlpDev& m, float (&h)[W], float* zs, int lstride, int cstride = 1) {
  constexpr int LS = W * W + W;
#pragma unroll
  for (int l = 0; l < ODE_MAXL; ++l) {
    if (l < m.L) {
      float o[W];
#pragma unroll
      for (int n = 0; n < W; ++n) o[n] = c_ode_w[l * LS + W * W + n];
      const int K = m.dims[l];
#pragma unroll
      for (int k = 0; k < W; ++k) {
        if (k < K) {
          const float hk = h[k];
#pragma unroll
          for (int n = 0; n < W; ++n) o[n] = fmaf(c_ode_w[l * LS + k * W + n], hk, o[n]);
        }
      }
      const int act = m.act[l];
#pragma unroll
      for (int n = 0; n < W; ++n) h[n] = ode_act(act, o[n]);
      if (KEEP && l + 1 < m.L) {
        float* z = zs + (size_t)(l + 1) * lstride;
#pragma unroll
        for (int n = 0; n < W; ++n) z[(size_t)n * cstride] = h[n];
      }
    }
  }
}

// Static per-edge data of a CTA, cached in shared memory once per launch: (src, dst) and the columns of the phi input that
// do not change between right-hand sides (static h_t, static h_s, pos_s - pos_t).
struct OdeEdgeCache {
  int* es;        // [max_edges] source node
  int* et;        // [max_edges] destination node
  float* stat;    // [max_edges][nstat], nstat = 2 dhs + dpos
  int nstat;
};

__device__ __forceinline__ void ode_fill_cache(const OdeArgs& a, const OdeEdgeCache& c, int k0, int k1, int tid, int nthreads) {
  const int ds = a.dhs + a.dpos;
  for (int k = k0 + tid; k < k1; k += nthreads) {
    const int s = a.src[k], t = a.dst[k];
    c.es[k - k0] = s;
    c.et[k - k0] = t;
    float* st = c.stat + (size_t)(k - k0) * c.nstat;
    for (int j = 0; j < a.dhs; ++j) {
      st[j] = a.snode[(size_t)t * ds + j];
      st[a.dhs + j] = a.snode[(size_t)s * ds + j];
    }
    for (int j = 0; j < a.dpos; ++j) st[2 * a.dhs + j] = a.snode[(size_t)s * ds + a.dhs + j] - a.snode[(size_t)t * ds + a.dhs + j];
  }
}

// the phi input of edge (s -> t):  [h_t; h_s; pos_s - pos_t],  h = [u (dx); static columns]   (layers.jl:104-106);
// `uin` is the [N][dx] stage input (shared memory in the forward kernel, the saved trajectory in the adjoint)
template <int W>
__device__ __forceinline__ void ode_input(const OdeArgs& a, const float* __restrict__ uin, int s, int t, const float* __restrict__ st,
                                          float (&in)[W]) {
  const int dh = a.dx + a.dhs;
#pragma unroll
  for (int c = 0; c < W; ++c) {
    float v = 0.f;
    if (c < ODE_MAXIN && c < a.din) {
      if (c < 2 * dh) {
        const int node = c < dh ? t : s, f = c < dh ? c : c - dh;
        v = f < a.dx ? uin[(size_t)node * a.dx + f] : st[(c < dh ? 0 : a.dhs) + (f - a.dx)];
      } else {
        v = st[2 * a.dhs + (c - 2 * dh)];
      }
    }
    in[c] = v;
  }
}

// Forward: every CTA keeps the whole current stage input in its own shared memory (double buffered); the owner of a node
// stores its next stage input into all CTAs' buffers through distributed shared memory, so a right-hand side touches global
// memory only to record the trajectory, and the one cluster barrier per right-hand side orders the DSMEM stores.
template <int W>
__global__ void __launch_bounds__(ODE_THREADS, 1) edgeconv_ode_fwd_kernel(const __grid_constant__ OdeArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float sm[];
  constexpr int LS = W * W + W;
  const int tid = threadIdx.x, cta = blockIdx.x;
  const int n0 = (int)((long long)a.N * cta / ODE_CTAS), n1 = (int)((long long)a.N * (cta + 1) / ODE_CTAS);
  const int k0 = a.rowptr[n0], k1 = a.rowptr[n1];
  const int dx = a.dx, nown = (n1 - n0) * dx, nd = a.N * dx;
  const int max_own = ((a.N + ODE_CTAS - 1) / ODE_CTAS + 1) * dx;
  float* msg = sm;                                  // [edges of this CTA][dx]
  float* ucur = msg + (((size_t)a.max_edges * dx + 3) & ~size_t(3));   // [2][N][dx] stage input, double buffered
  float* kst = ucur + 2 * (((size_t)nd + 3) & ~size_t(3));             // [S][own][dx] stage derivatives of the owned nodes
  float* u0 = kst + (size_t)a.S * max_own;           // [own][dx] state at the start of the step
  OdeEdgeCache ec;
  ec.nstat = 2 * a.dhs + a.dpos;
  ec.stat = u0 + max_own;
  ec.es = reinterpret_cast<int*>(ec.stat + (((size_t)a.max_edges * ec.nstat + 3) & ~size_t(3)));
  ec.et = ec.es + a.max_edges;
  int* rp = ec.et + a.max_edges;                     // [own + 1] row pointers of the owned nodes
  ode_fill_cache(a, ec, k0, k1, tid, ODE_THREADS);
  for (int i = tid; i <= n1 - n0; i += ODE_THREADS) rp[i] = a.rowptr[n0 + i];
  const size_t ubuf = ((size_t)nd + 3) & ~size_t(3);
  for (int i = tid; i < nd; i += ODE_THREADS) ucur[i] = a.u[i];   // every CTA reads the whole initial state
  for (int i = tid; i < nown; i += ODE_THREADS) {
    u0[i] = a.u[n0 * dx + i];
    a.traj[n0 * dx + i] = u0[i];                    // stage 0 input of step 0 = u
  }
  cluster.sync();
  int cur = 0;
  for (int step = 0; step < a.n_steps; ++step) {
    float* tr = a.traj + (size_t)step * a.S * nd;
    for (int s = 0; s < a.S; ++s) {
      const float* uin = ucur + (size_t)cur * ubuf;
      // ---- edges: messages ----
      for (int e = tid; e < k1 - k0; e += ODE_THREADS) {
        float h[W];
        ode_input<W>(a, uin, ec.es[e], ec.et[e], ec.stat + (size_t)e * ec.nstat, h);
        ode_mlp<W, false>(a.mlp, h, nullptr, 0);
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c < dx) msg[(size_t)e * dx + c] = h[c];
      }
      __syncthreads();
      // ---- nodes: ordered aggregation, stage derivative, next stage's input to every CTA ----
      float* unext = ucur + (size_t)(cur ^ 1) * ubuf;
      for (int i = tid; i < nown; i += ODE_THREADS) {
        const int nl = i / dx, c = i - nl * dx;
        const int r0 = rp[nl] - k0, r1 = rp[nl + 1] - k0;
        float acc = 0.f;
        for (int k = r0; k < r1; ++k) acc = __fadd_rn(acc, msg[(size_t)k * dx + c]);
        if (a.aggr == NGPDE_AGGR_MEAN && r1 > r0) acc = __fdiv_rn(acc, (float)(r1 - r0));
        kst[(size_t)s * max_own + i] = acc;
        float nxt = u0[i];
        const bool last = s + 1 == a.S;
        for (int j = 0; j <= s; ++j) {
          const float cf = a.dt * (last ? a.b[j] : a.a[s + 1][j]);
          if (cf != 0.f) nxt = fmaf(cf, kst[(size_t)j * max_own + i], nxt);
        }
        const int gi = n0 * dx + i;
        if (!last) {
          tr[(size_t)(s + 1) * nd + gi] = nxt;
        } else {
          if (step + 1 < a.n_steps) a.traj[(size_t)(step + 1) * a.S * nd + gi] = nxt;
          a.u[gi] = nxt;
          u0[i] = nxt;
        }
#pragma unroll
        for (int r = 0; r < ODE_CTAS; ++r) cluster.map_shared_rank(unext, r)[gi] = nxt;
      }
      cluster.sync();
      cur ^= 1;
    }
  }
}

// ---- adjoint ----
struct OdeBwdSmem {
  int off_kb, off_dte, off_z, off_g, off_stat, off_es, off_et, nstat, te, floats;
};

inline OdeBwdSmem ode_bwd_smem(const MlpDev& m, int W, int dx, int max_nodes, int max_edges, int nstat) {
  OdeBwdSmem s;
  int off = 0;
  s.nstat = nstat;
  s.off_stat = off; off += (max_edges * nstat + 3) & ~3;
  s.off_es = off;   off += (max_edges + 3) & ~3;
  s.off_et = off;   off += (max_edges + 3) & ~3;
  s.off_kb = off;  off += (max_nodes * dx + 3) & ~3;     // stage cotangent of the owned nodes (already / deg)
  s.off_dte = off; off += (max_edges * dx + 3) & ~3;     // destination-side input cotangent per edge of this CTA
  s.te = W <= 16 ? 256 : 128;                            // edges per tile
  s.off_z = off;   off += m.L * W * s.te;                // layer inputs Z_0 .. Z_{L-1} of the tile, [layer][column][edge]
  s.off_g = off;   off += m.L * W * s.te;                // pre-activation cotangents of the tile, same layout
  s.floats = off;
  return s;
}

// Adjoint tiles are stored COLUMN-major ([layer][column][edge]): the edge threads' stores are conflict-free (lane = edge) and
// the parameter-gradient pass reads four consecutive edges of a column with one 128-bit load.  That pass: warp <-> (layer,
// block of 4 input rows), lane <-> (block of 4 output columns, edge slice): every thread accumulates a 4 x 4 block of dW
// (+ 4 bias sums when its rows start at 0) over its slice's edges in ascending order, in registers, for the whole kernel;
// the slices are combined at the end by a fixed shuffle tree -- deterministic, no atomics.
template <int W>
__global__ void __launch_bounds__(ODE_THREADS, 1) edgeconv_ode_bwd_kernel(const __grid_constant__ OdeArgs a, const OdeBwdSmem L) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float sm[];
  constexpr int LS = W * W + W;
  constexpr int NB = W / 4;              // 4-wide blocks per matrix side
  constexpr int SL = 32 / NB;            // edge slices (lanes per output block)
  constexpr int PAIRS = ODE_MAXL * NB;   // (layer, row block) pairs
  constexpr int PPW = PAIRS / (ODE_THREADS / 32);  // pairs per warp: 1 (W = 16) or 2 (W = 32)
  static_assert(PPW >= 1 && PPW * (ODE_THREADS / 32) == PAIRS, "warp <-> (layer, row block) mapping");
  float* kb = sm + L.off_kb;
  float* dte = sm + L.off_dte;
  float* Z = sm + L.off_z;
  float* G = sm + L.off_g;
  const MlpDev& m = a.mlp;
  const int tid = threadIdx.x, cta = blockIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = (int)((long long)a.N * cta / ODE_CTAS), n1 = (int)((long long)a.N * (cta + 1) / ODE_CTAS);
  const int k0 = a.rowptr[n0], k1 = a.rowptr[n1];
  const int dx = a.dx, nl = m.L, TE = L.te, lstride = W * L.te;
  const size_t nd = (size_t)a.N * dx;
  OdeEdgeCache ec;
  ec.nstat = L.nstat;
  ec.stat = sm + L.off_stat;
  ec.es = reinterpret_cast<int*>(sm + L.off_es);
  ec.et = reinterpret_cast<int*>(sm + L.off_et);
  ode_fill_cache(a, ec, k0, k1, tid, ODE_THREADS);
  const int nb = lane / SL, slice = lane % SL;
  float dw[PPW][16], db[PPW][4];
#pragma unroll
  for (int r = 0; r < PPW; ++r) {
#pragma unroll
    for (int j = 0; j < 16; ++j) dw[r][j] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) db[r][j] = 0.f;
  }
  int buf = 0;
  __syncthreads();
  for (int step = a.n_steps - 1; step >= 0; --step) {
    const float* tr = a.traj + (size_t)step * a.S * nd;
    for (int s = a.S - 1; s >= 0; --s) {
      const float* uin = tr + (size_t)s * nd;
      // ---- stage cotangent of the owned nodes: kbar_s = dt b_s lam + dt sum_{i > s} a_is ubar_i ----
      for (int i = n0 * dx + tid; i < n1 * dx; i += ODE_THREADS) {
        float v = a.dt * a.b[s] * a.lam[i];
        for (int j = s + 1; j < a.S; ++j) {
          const float cf = a.dt * a.a[j][s];
          if (cf != 0.f) v = fmaf(cf, a.kbuf[(size_t)j * nd + i], v);
        }
        const int node = i / dx;
        const int deg = a.rowptr[node + 1] - a.rowptr[node];
        if (a.aggr == NGPDE_AGGR_MEAN && deg > 0) v = __fdiv_rn(v, (float)deg);
        kb[i - n0 * dx] = v;
      }
      __syncthreads();
      float* desrc = a.desrc + (size_t)buf * a.E * dx;
      // ---- edge tiles: recompute phi, back-propagate, parameter gradient ----
      for (int t0 = k0; t0 < k1; t0 += TE) {
        const int ne = min(TE, k1 - t0);
        if (tid < ne) {
          const int k = t0 + tid, sidx = ec.es[k - k0], didx = ec.et[k - k0];
          float h[W];
          ode_input<W>(a, uin, sidx, didx, ec.stat + (size_t)(k - k0) * ec.nstat, h);  // the trajectory was written by the forward launch
#pragma unroll
          for (int c = 0; c < W; ++c) Z[(size_t)c * TE + tid] = h[c];
          ode_mlp<W, true>(m, h, Z + tid, lstride, TE);
          // cotangent of the output: the (scaled) stage cotangent of the destination
          float g[W];
#pragma unroll
          for (int n = 0; n < W; ++n) g[n] = (n < dx) ? kb[(size_t)(didx - n0) * dx + (n < dx ? n : 0)] : 0.f;
#pragma unroll
          for (int l = ODE_MAXL - 1; l >= 0; --l) {
            if (l < nl) {
              // through the activation: layer l's output is h (last layer) or the kept input of layer l + 1
              const int act = m.act[l];
              if (act != NGPDE_ACT_IDENTITY) {
                const float* y = Z + (size_t)(l + 1) * lstride + tid;
#pragma unroll
                for (int n = 0; n < W; ++n) g[n] *= act_grad_y(act, (l == nl - 1) ? h[n] : y[(size_t)n * TE]);
              }
              float* gt = G + (size_t)l * lstride + tid;
#pragma unroll
              for (int n = 0; n < W; ++n) gt[(size_t)n * TE] = g[n];
              float dh[W];
#pragma unroll
              for (int k2 = 0; k2 < W; ++k2) {
                float acc = 0.f;
#pragma unroll
                for (int n = 0; n < W; ++n) acc = fmaf(c_ode_w[l * LS + k2 * W + n], g[n], acc);
                dh[k2] = acc;
              }
#pragma unroll
              for (int k2 = 0; k2 < W; ++k2) g[k2] = dh[k2];
            }
          }
          // g now holds d/d(input): [h_t (dx + dhs); h_s (dx + dhs); dpos]
          const int dh1 = dx + a.dhs;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c < dx) {
              float vs = 0.f;
#pragma unroll
              for (int j = 0; j < W; ++j)
                if (j == dh1 + c) vs = g[j];
              dte[(size_t)(k - k0) * dx + c] = g[c];
              desrc[(size_t)k * dx + c] = vs;
            }
          }
        } else if (tid < TE) {
          for (int l = 0; l < nl; ++l) {  // rows beyond the tile contribute nothing to the parameter gradient
#pragma unroll
            for (int n = 0; n < W; ++n) {
              G[(size_t)l * lstride + (size_t)n * TE + tid] = 0.f;
              Z[(size_t)l * lstride + (size_t)n * TE + tid] = 0.f;
            }
          }
        }
        __syncthreads();
        // ---- parameter gradient of the tile ----
#pragma unroll
        for (int r = 0; r < PPW; ++r) {
          const int pair = warp * PPW + r, l = pair / NB, kbk = pair - l * NB;
          if (l < nl) {
            const float* zc = Z + (size_t)l * lstride + (size_t)(4 * kbk) * TE;
            const float* gc = G + (size_t)l * lstride + (size_t)(4 * nb) * TE;
            for (int ch = slice; ch < TE / 4; ch += SL) {  // 4 consecutive edges per trip, slices interleaved
              float4 z4[4], g4[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                z4[i] = *reinterpret_cast<const float4*>(zc + (size_t)i * TE + 4 * ch);
                g4[i] = *reinterpret_cast<const float4*>(gc + (size_t)i * TE + 4 * ch);
              }
#pragma unroll
              for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float acc = dw[r][4 * i + j];
                  acc = fmaf(z4[i].x, g4[j].x, acc);
                  acc = fmaf(z4[i].y, g4[j].y, acc);
                  acc = fmaf(z4[i].z, g4[j].z, acc);
                  acc = fmaf(z4[i].w, g4[j].w, acc);
                  dw[r][4 * i + j] = acc;
                }
              if (kbk == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) db[r][j] += ((g4[j].x + g4[j].y) + g4[j].z) + g4[j].w;
              }
            }
          }
        }
        __syncthreads();
      }
      __threadfence();
      cluster.sync();  // every CTA's source-side cotangents of this stage are visible
      // ---- ubar_s of the owned nodes: destination side (own edges, CSR order) + source side (transpose order) ----
      for (int i = n0 * dx + tid; i < n1 * dx; i += ODE_THREADS) {
        const int node = i / dx, c = i - node * dx;
        float acc = 0.f;
        for (int k = a.rowptr[node]; k < a.rowptr[node + 1]; ++k) acc += dte[(size_t)(k - k0) * dx + c];
        for (int q = a.tptr[node]; q < a.tptr[node + 1]; ++q) acc += __ldcg(desrc + (size_t)a.tpos[q] * dx + c);
        a.kbuf[(size_t)s * nd + i] = acc;
      }
      buf ^= 1;
      __syncthreads();
    }
    // lam <- lam + sum_s ubar_s  (own nodes)
    for (int i = n0 * dx + tid; i < n1 * dx; i += ODE_THREADS) {
      float v = a.lam[i];
      for (int s = 0; s < a.S; ++s) v += a.kbuf[(size_t)s * nd + i];
      a.lam[i] = v;
    }
    __syncthreads();
  }
  // ---- parameter gradient: slices combined by a fixed shuffle tree, CTA partials summed in CTA order by CTA 0 ----
  for (int p = tid; p < m.n_params; p += ODE_THREADS) a.dpart[(size_t)cta * m.n_params + p] = 0.f;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < PPW; ++r) {
    const int pair = warp * PPW + r, l = pair / NB, kbk = pair - l * NB;
#pragma unroll
    for (int j = 0; j < 20; ++j) {
      float v = j < 16 ? dw[r][j < 16 ? j : 0] : db[r][j >= 16 ? j - 16 : 0];
#pragma unroll
      for (int off = SL / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (slice == 0 && l < nl) {
        const int K = m.dims[l], N = m.dims[l + 1];
        if (j < 16) {
          const int k = 4 * kbk + j / 4, n = 4 * nb + (j & 3);
          if (k < K && n < N) a.dpart[(size_t)cta * m.n_params + m.w_off[l] + k * N + n] = v;
        } else if (kbk == 0 && m.b_off[l] >= 0) {
          const int n = 4 * nb + (j - 16);
          if (n < N) a.dpart[(size_t)cta * m.n_params + m.b_off[l] + n] = v;
        }
      }
    }
  }
  __threadfence();
  cluster.sync();
  if (cta == 0) {
    for (int p = tid; p < m.n_params; p += ODE_THREADS) {
      float s = 0.f;
      for (int c = 0; c < ODE_CTAS; ++c) s += __ldcg(a.dpart + (size_t)c * m.n_params + p);
      a.dparams[p] = s;
    }
  }
}

int make_ode_mlp(const ngpde_mlp& m, MlpDev* out) {
  out->L = m.n_layers;
  int off = 0;
  for (int l = 0; l <= m.n_layers; ++l) out->dims[l] = m.dims[l];
  for (int l = 0; l < m.n_layers; ++l) {
    out->act[l] = m.act[l];
    out->w_off[l] = off;
    off += m.dims[l] * m.dims[l + 1];
    if (m.has_bias[l]) { out->b_off[l] = off; off += m.dims[l + 1]; } else out->b_off[l] = -1;
  }
  out->n_params = off;
  return NGPDE_OK;
}

// compile-time layer width the kernels run with: 16 when every layer (and the input) fits, else 32
int ode_width(const MlpDev& m) {
  int w = 0;
  for (int l = 0; l <= m.L; ++l) w = std::max(w, m.dims[l]);
  return w <= 16 ? 16 : 32;
}

struct OdePlan {
  OdeArgs a{};
  size_t off_kbuf = 0, off_desrc = 0, off_dpart = 0, off_wpad = 0, total = 0;
  int max_nodes = 0;
};

int ode_plan(const ngpde_graph* g, const ngpde_conv_desc* d, const ngpde_rk_tableau* tab, OdePlan* P) {
  NGPDE_REQUIRE(g && d && tab, "null argument");
  NGPDE_REQUIRE(d->family == NGPDE_EXPLICIT_EDGE_CONV, "the persistent ODE kernel integrates ExplicitEdgeConv right-hand sides");
  NGPDE_REQUIRE(d->aggr == NGPDE_AGGR_MEAN || d->aggr == NGPDE_AGGR_SUM, "persistent ODE kernel: aggr must be + or mean");
  NGPDE_REQUIRE(d->phi.n_layers >= 1 && d->phi.n_layers <= 4, "persistent ODE kernel: phi must have 1..4 Dense layers");
  NGPDE_REQUIRE(d->dx >= 1 && d->dx <= 4, "persistent ODE kernel: state width must be 1..4");
  const int din = 2 * (d->dx + d->dhs) + d->dpos;
  NGPDE_REQUIRE(din == d->phi.dims[0] && din <= ODE_MAXIN, "persistent ODE kernel: phi input %d (limit %d)", din, ODE_MAXIN);
  for (int l = 1; l <= d->phi.n_layers; ++l)
    NGPDE_REQUIRE(d->phi.dims[l] <= ODE_MAXW, "persistent ODE kernel: phi layer %d is %d wide (limit %d)", l, d->phi.dims[l], ODE_MAXW);
  NGPDE_REQUIRE(d->phi.dims[d->phi.n_layers] == d->dx, "an ODE right-hand side must return the state's width");
  for (int l = 0; l < d->phi.n_layers; ++l)
    NGPDE_REQUIRE(act_grad_from_y(d->phi.act[l]), "persistent ODE kernel: swish / gelu are not supported");
  NGPDE_REQUIRE(tab->n_stages >= 1 && tab->n_stages <= ODE_MAXS, "tableau: 1..%d stages", ODE_MAXS);
  NGPDE_REQUIRE(g->N >= ODE_CTAS && g->N <= (1 << 16), "persistent ODE kernel: %d <= N <= 65536 nodes", ODE_CTAS);
  OdeArgs& a = P->a;
  a.rowptr = g->rowptr; a.src = g->src; a.dst = g->dst; a.tptr = g->tptr; a.tpos = g->tpos;
  a.N = (int)g->N; a.E = (int)g->E;
  a.dx = d->dx; a.dhs = d->dhs; a.dpos = d->dpos; a.din = din; a.aggr = d->aggr;
  make_ode_mlp(d->phi, &a.mlp);
  NGPDE_REQUIRE(a.mlp.n_params <= 4 * ODE_THREADS * 4, "persistent ODE kernel: too many parameters");
  a.S = tab->n_stages;
  std::memcpy(a.a, tab->a, sizeof(a.a));
  std::memcpy(a.b, tab->b, sizeof(a.b));
  P->max_nodes = (int)((g->N + ODE_CTAS - 1) / ODE_CTAS) + 1;
  size_t off = 0;
  const size_t nd = (size_t)g->N * d->dx;
  P->off_kbuf = off;  off = (off + sizeof(float) * ODE_MAXS * nd + 255) & ~size_t(255);
  P->off_desrc = off; off = (off + sizeof(float) * 2 * (size_t)g->E * d->dx + 255) & ~size_t(255);
  P->off_dpart = off; off = (off + sizeof(float) * ODE_CTAS * a.mlp.n_params + 255) & ~size_t(255);
  P->off_wpad = off;  off = (off + sizeof(c_ode_w) + 255) & ~size_t(255);
  P->total = off + 256;
  return NGPDE_OK;
}

// largest per-CTA edge count: rowptr lives on the device -> read the ODE_CTAS + 1 boundary entries
int ode_max_edges(ngpde_graph* g, cudaStream_t st, int* out) {
  if (g->ode_max_edges >= 0) {
    *out = g->ode_max_edges;
    return NGPDE_OK;
  }
  int h[ODE_CTAS + 1];
  for (int c = 0; c <= ODE_CTAS; ++c) {
    const int n = (int)((long long)g->N * c / ODE_CTAS);
    NGPDE_CUDA_TRY(cudaMemcpyAsync(&h[c], g->rowptr + n, sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  NGPDE_CUDA_TRY(cudaStreamSynchronize(st));
  int m = 0;
  for (int c = 0; c < ODE_CTAS; ++c) m = std::max(m, h[c + 1] - h[c]);
  g->ode_max_edges = m;
  *out = m;
  return NGPDE_OK;
}

// padded parameters -> constant bank (device-to-device, in stream order)
int ode_stage_params(const OdeArgs& a, int W, float* staging, cudaStream_t st) {
  ode_pad_params_kernel<<<8, 256, 0, st>>>(a.mlp, a.params, W, staging);
  NGPDE_CUDA_TRY(cudaGetLastError());
  NGPDE_CUDA_TRY(cudaMemcpyToSymbolAsync(c_ode_w, staging, sizeof(float) * (size_t)a.mlp.L * (W * W + W), 0, cudaMemcpyDeviceToDevice, st));
  return NGPDE_OK;
}

template <class K, class... Args>
int launch_cluster(K kernel, size_t smem, cudaStream_t st, Args... args) {
  NGPDE_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ODE_CTAS);
  cfg.blockDim = dim3(ODE_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ODE_CTAS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  NGPDE_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, args...));
  return NGPDE_OK;
}

}  // namespace
}  // namespace ngpde

using namespace ngpde;

extern "C" size_t ngpde_edgeconv_ode_workspace_bytes(ngpde_graph_t g, const ngpde_conv_desc* desc, const ngpde_rk_tableau* tab) {
  OdePlan P;
  if (ode_plan(g, desc, tab, &P)) return 0;
  return P.total;
}

extern "C" int ngpde_edgeconv_ode_forward(ngpde_graph_t g, const ngpde_conv_desc* desc, const ngpde_rk_tableau* tab, float dt,
                                          int32_t n_steps, const float* phi_params, const float* snode, float* u, float* traj,
                                          void* workspace, size_t workspace_bytes, void* stream) {
  OdePlan P;
  if (int rc = ode_plan(g, desc, tab, &P)) return rc;
  NGPDE_REQUIRE(phi_params && u && traj && n_steps >= 1, "ode_forward: null argument / n_steps < 1");
  NGPDE_REQUIRE(desc->dhs + desc->dpos == 0 || snode, "ode_forward: snode is NULL");
  NGPDE_REQUIRE(workspace && workspace_bytes >= P.total, "ode_forward: workspace too small (%zu < %zu)", workspace_bytes, P.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  OdeArgs& a = P.a;
  if (int rc = ode_max_edges(g, st, &a.max_edges)) return rc;
  char* ws = static_cast<char*>(workspace);
  a.params = phi_params; a.snode = snode; a.dt = dt; a.n_steps = n_steps; a.u = u; a.traj = traj;
  a.kbuf = reinterpret_cast<float*>(ws + P.off_kbuf);
  const int W = ode_width(a.mlp);
  const int nd = a.N * a.dx, max_own = ((a.N + ODE_CTAS - 1) / ODE_CTAS + 1) * a.dx, nstat = 2 * a.dhs + a.dpos;
  const size_t fl = (((size_t)a.max_edges * a.dx + 3) & ~size_t(3)) + 2 * (((size_t)nd + 3) & ~size_t(3)) +
                    (size_t)a.S * max_own + max_own + (((size_t)a.max_edges * nstat + 3) & ~size_t(3)) + 2 * (size_t)a.max_edges +
                    (size_t)max_own + 8;
  const size_t smem = sizeof(float) * fl;
  NGPDE_REQUIRE(smem <= 200 * 1024, "ode_forward: the graph (%d nodes, %d edges per CTA) does not fit shared memory; use the "
                "CUDA-graph step path", a.N, a.max_edges);
  if (int rc = ode_stage_params(a, W, reinterpret_cast<float*>(ws + P.off_wpad), st)) return rc;
  return W == 16 ? launch_cluster(edgeconv_ode_fwd_kernel<16>, smem, st, a) : launch_cluster(edgeconv_ode_fwd_kernel<32>, smem, st, a);
}

extern "C" int ngpde_edgeconv_ode_adjoint(ngpde_graph_t g, const ngpde_conv_desc* desc, const ngpde_rk_tableau* tab, float dt,
                                          int32_t n_steps, const float* phi_params, const float* snode, const float* traj,
                                          float* lam, float* dphi_params, void* workspace, size_t workspace_bytes, void* stream) {
  OdePlan P;
  if (int rc = ode_plan(g, desc, tab, &P)) return rc;
  NGPDE_REQUIRE(phi_params && traj && lam && dphi_params && n_steps >= 1, "ode_adjoint: null argument / n_steps < 1");
  NGPDE_REQUIRE(desc->dhs + desc->dpos == 0 || snode, "ode_adjoint: snode is NULL");
  NGPDE_REQUIRE(workspace && workspace_bytes >= P.total, "ode_adjoint: workspace too small (%zu < %zu)", workspace_bytes, P.total);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  OdeArgs& a = P.a;
  if (int rc = ode_max_edges(g, st, &a.max_edges)) return rc;
  char* ws = static_cast<char*>(workspace);
  a.params = phi_params; a.snode = snode; a.dt = dt; a.n_steps = n_steps;
  a.traj = const_cast<float*>(traj); a.lam = lam; a.dparams = dphi_params;
  a.kbuf = reinterpret_cast<float*>(ws + P.off_kbuf);
  a.desrc = reinterpret_cast<float*>(ws + P.off_desrc);
  a.dpart = reinterpret_cast<float*>(ws + P.off_dpart);
  const int W = ode_width(a.mlp);
  const OdeBwdSmem L = ode_bwd_smem(a.mlp, W, a.dx, P.max_nodes, a.max_edges, 2 * a.dhs + a.dpos);
  const size_t smem = sizeof(float) * (size_t)L.floats + 64;
  NGPDE_REQUIRE(smem <= 220 * 1024, "ode_adjoint: the tile buffers need %zu bytes of shared memory", smem);
  if (int rc = ode_stage_params(a, W, reinterpret_cast<float*>(ws + P.off_wpad), st)) return rc;
  return W == 16 ? launch_cluster(edgeconv_ode_bwd_kernel<16>, smem, st, a, L) : launch_cluster(edgeconv_ode_bwd_kernel<32>, smem, st, a, L);
}
