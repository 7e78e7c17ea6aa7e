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
long long* tc_get_debug_buffer();  // ngpde_tc.cu (ngpde_debug_buffer)
namespace {

constexpr int ODE_CTAS = 8;       // portable cluster size (the FFMA kernels; lower bound on N)
constexpr int ODE_MAX_CTAS = 16;  // non-portable cluster size the tensor-core kernels ask for when the device grants it
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
  int ncta;                   // CTAs in the cluster (8 or 16)
  long long* dbg;             // optional phase cycle sums of CTA 0, thread 0 (ngpde_debug_buffer; tools/ode_bench.py)
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

// phase profiling (developer aid): thread 0 of CTA 0 sums clock64 differences per phase and writes them at the end
#define ODE_PROF_DECL(n) long long prof_t = 0, prof_t2 = 0, prof_acc[n] = {}; (void)prof_t; (void)prof_t2; (void)prof_acc
#define ODE_PROF_START() do { if (a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) prof_t = clock64(); } while (0)
#define ODE_PROF(i) do { if (a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { const long long t_ = clock64(); prof_acc[i] += t_ - prof_t; prof_t = t_; } } while (0)
#define ODE_PROF_START2() do { if (a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) prof_t2 = clock64(); } while (0)
#define ODE_PROF2(i) do { if (a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { const long long t_ = clock64(); prof_acc[i] += t_ - prof_t2; prof_t2 = t_; } } while (0)
#define ODE_PROF_END(base, n) do { if (a.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) for (int i_ = 0; i_ < n; ++i_) a.dbg[(base) + i_] = prof_acc[i_]; } while (0)

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
  const int n0 = (int)((long long)a.N * cta / a.ncta), n1 = (int)((long long)a.N * (cta + 1) / a.ncta);
  const int k0 = a.rowptr[n0], k1 = a.rowptr[n1];
  const int dx = a.dx, nown = (n1 - n0) * dx, nd = a.N * dx;
  const int max_own = ((a.N + a.ncta - 1) / a.ncta + 1) * dx;
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
  ODE_PROF_DECL(4);
  for (int step = 0; step < a.n_steps; ++step) {
    float* tr = a.traj + (size_t)step * a.S * nd;
    for (int s = 0; s < a.S; ++s) {
      const float* uin = ucur + (size_t)cur * ubuf;
      ODE_PROF_START();
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
      ODE_PROF(0);
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
        for (int r = 0; r < a.ncta; ++r) cluster.map_shared_rank(unext, r)[gi] = nxt;
      }
      ODE_PROF(1);
      cluster.sync();
      ODE_PROF(2);
      cur ^= 1;
    }
  }
  ODE_PROF_END(0, 3);
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
  const int n0 = (int)((long long)a.N * cta / a.ncta), n1 = (int)((long long)a.N * (cta + 1) / a.ncta);
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
  ODE_PROF_DECL(8);
  for (int step = a.n_steps - 1; step >= 0; --step) {
    const float* tr = a.traj + (size_t)step * a.S * nd;
    for (int s = a.S - 1; s >= 0; --s) {
      const float* uin = tr + (size_t)s * nd;
      ODE_PROF_START();
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
      ODE_PROF(0);
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
        ODE_PROF(1);
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
        ODE_PROF(2);
      }
      __threadfence();
      cluster.sync();  // every CTA's source-side cotangents of this stage are visible
      ODE_PROF(3);
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
      ODE_PROF(4);
    }
    ODE_PROF_START();
    // lam <- lam + sum_s ubar_s  (own nodes)
    for (int i = n0 * dx + tid; i < n1 * dx; i += ODE_THREADS) {
      float v = a.lam[i];
      for (int s = 0; s < a.S; ++s) v += a.kbuf[(size_t)s * nd + i];
      a.lam[i] = v;
    }
    __syncthreads();
    ODE_PROF(5);
  }
  ODE_PROF_END(8, 6);
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
      for (int c = 0; c < a.ncta; ++c) s += __ldcg(a.dpart + (size_t)c * m.n_params + p);
      a.dparams[p] = s;
    }
  }
}

// ================================================================================================
// Tensor-core form of both kernels for phi layers <= 16 wide (C1: 4 => 16 => 16 => 1).
//
// The FFMA kernels above spend a right-hand side streaming weights: one thread per edge needs every weight once, so a CTA of
// 16 warps issues 16 x 816 weight operands per right-hand side whether they come from shared memory (load pipe) or from the
// constant bank (its 2 KB first-level cache thrashes on the 3.2 KB image): measured 17k cycles per forward right-hand side
// and 47k per 256-edge adjoint tile (tools/ode_bench.py).  Here a warp owns 16 edges at a time and every Dense layer is
// `mma.sync.m16n8k8` on TF32 operands with the 3xTF32 split (a b ~ a_lo b_hi + a_hi b_lo + a_hi b_hi, FP32 accumulate:
// float32-accurate to ~1e-6, the north-star tolerance is 1e-5 per call / 1e-4 on the trajectory).  The accumulator fragment
// of one layer IS the A fragment of the next after a relabelling of the K index (lane (g, t) holds columns 8j + 2t and
// 8j + 2t + 1 of accumulator tile j; as A columns t and t + 4 of K-step j), which is folded into the weight-fragment
// images built once per launch in shared memory -- no shuffles, no shared-memory round trip between layers.
// A 16-row tile is too small for tcgen05 (M = 128, one issuing thread, TMEM round trip): the legacy warp-level MMA is the
// right tensor-core instruction for 16 x 16 layers whose time is latency, not throughput.
//
// Adjoint, per 16-edge tile: recompute (fragments of every layer input stay in registers), then per layer backwards
// dH = G W' (same trick, transposed weight images) and dW += Z' G, whose operands need the edge index on the K side: both
// are transposed through a warp-private 2.5 KB shared-memory patch.  Every warp accumulates dW / db of ITS edges in
// registers for the whole launch; warps are added in warp order, CTAs in CTA order at the end (deterministic).  The
// source-side input cotangents go straight into the owner CTA's shared memory (DSMEM) at their transpose-order slot;
// state, stage cotangents and lam of the owned nodes live in shared memory: global memory is touched only to prefetch the
// next stage input (cp.async) -- one cluster barrier + one CTA barrier per right-hand side.
// ================================================================================================
constexpr int ODE_PS = 20;  // patch row stride (floats): conflict-free for the fragment stores AND the transposed fragment loads

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  lo = __float_as_uint(x - __uint_as_float(hi));  // the MMA reads its upper 19 bits: |lo| <= 2^-11 |x|, so the cut is <= 2^-22 |x|
}
// c += a * b for one (K-step, N-tile): a given as FP32 values in A-fragment order, b as the prepared {b0_hi, b1_hi, b0_lo, b1_lo}
__device__ __forceinline__ void mma_3x(float (&c)[4], const uint32_t (&ahi)[4], const uint32_t (&alo)[4], const float4 b) {
  mma_tf32(c, alo, __float_as_uint(b.x), __float_as_uint(b.y));
  mma_tf32(c, ahi, __float_as_uint(b.z), __float_as_uint(b.w));
  mma_tf32(c, ahi, __float_as_uint(b.x), __float_as_uint(b.y));
}
// accumulator-layout tile (rows g, g + 8; columns 2t, 2t + 1) -> A fragment of the K-step with the relabelled K index
__device__ __forceinline__ void frag_c_to_a(const float (&c)[4], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
  split_tf32(c[0], hi[0], lo[0]);
  split_tf32(c[2], hi[1], lo[1]);
  split_tf32(c[1], hi[2], lo[2]);
  split_tf32(c[3], hi[3], lo[3]);
}

// shared-memory carve-up shared by host and device (float offsets; every array 16-byte aligned)
struct OdeMmaSmem {
  int off_wf, off_wt, off_bias;       // weight-fragment images [L][2][2][32] float4 (forward, transposed), biases [L][16]
  int off_stat, off_es, off_et, off_rp, off_ucur, ubuf;
  // forward
  int off_msg, off_kst, off_u0, max_own;
  // adjoint
  int off_tp, off_eown, off_kb, off_ub, off_lam, off_dte, off_dsrc, dsrc_stride, off_patch, off_acc;
  int nstat, floats;
};

inline int up4(int x) { return (x + 3) & ~3; }

inline OdeMmaSmem ode_mma_smem(const OdeArgs& a, bool adjoint, int max_tedges) {
  OdeMmaSmem s{};
  const int L = a.mlp.L, nd = a.N * a.dx, own_nodes = (a.N + a.ncta - 1) / a.ncta + 1;
  int off = 0;
  s.nstat = 2 * a.dhs + a.dpos;
  s.max_own = own_nodes * a.dx;
  s.off_wf = off;   off += L * 4 * 32 * 4;
  s.off_wt = off;   off += adjoint ? L * 4 * 32 * 4 : 0;
  s.off_bias = off; off += L * 16;
  s.off_stat = off; off += up4(a.max_edges * s.nstat);
  s.off_es = off;   off += up4(a.max_edges);
  s.off_et = off;   off += up4(a.max_edges);
  s.off_rp = off;   off += up4(own_nodes + 1);
  s.ubuf = up4(nd);
  s.off_ucur = off; off += 2 * s.ubuf;
  if (!adjoint) {
    s.off_msg = off; off += up4(a.max_edges * a.dx);
    s.off_kst = off; off += up4(a.S * s.max_own);
    s.off_u0 = off;  off += up4(s.max_own);
  } else {
    s.off_tp = off;    off += up4(own_nodes + 1);
    s.off_eown = off;  off += up4(a.max_edges);
    s.off_kb = off;    off += up4(s.max_own);
    s.off_ub = off;    off += up4(a.S * s.max_own);
    s.off_lam = off;   off += up4(s.max_own);
    s.off_dte = off;   off += up4(a.max_edges * a.dx);
    s.dsrc_stride = up4(max_tedges * a.dx);
    s.off_dsrc = off;  off += 2 * s.dsrc_stride;
    s.off_patch = off; off += (ODE_THREADS / 32) * 2 * 16 * ODE_PS;
    s.off_acc = off;   off += up4(a.mlp.n_params);
  }
  s.floats = off;
  return s;
}

// Weight-fragment images.  Forward image of layer l, K-step ks, N-tile nt, lane (g, t):
//   {hi W[8ks+2t][8nt+g], hi W[8ks+2t+1][8nt+g], lo .., lo ..}      (W = [in][out], zero outside the layer's real shape)
// transposed image (input gradient dH = G W'), K-step j over the layer's outputs, N-tile kt over its inputs:
//   {hi W[8kt+g][8j+2t], hi W[8kt+g][8j+2t+1], lo .., lo ..}
__device__ __forceinline__ void ode_fill_weight_images(const MlpDev& m, const float* __restrict__ params, float4* wf, float4* wt,
                                                       float* bias, int tid, int nthreads) {
  for (int i = tid; i < m.L * 128; i += nthreads) {
    const int l = i >> 7, ks = (i >> 6) & 1, nt = (i >> 5) & 1, lane = i & 31, g = lane >> 2, t = lane & 3;
    const int K = m.dims[l], N = m.dims[l + 1];
    const float* w = params + m.w_off[l];
    {
      const int k0 = 8 * ks + 2 * t, n = 8 * nt + g;
      const float v0 = (k0 < K && n < N) ? w[k0 * N + n] : 0.f, v1 = (k0 + 1 < K && n < N) ? w[(k0 + 1) * N + n] : 0.f;
      uint32_t h0, l0, h1, l1;
      split_tf32(v0, h0, l0);
      split_tf32(v1, h1, l1);
      wf[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
    }
    if (wt != nullptr) {  // here `ks` indexes the output tile j and `nt` the input tile kt
      const int n0 = 8 * ks + 2 * t, k = 8 * nt + g;
      const float v0 = (k < K && n0 < N) ? w[k * N + n0] : 0.f, v1 = (k < K && n0 + 1 < N) ? w[k * N + n0 + 1] : 0.f;
      uint32_t h0, l0, h1, l1;
      split_tf32(v0, h0, l0);
      split_tf32(v1, h1, l1);
      wt[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
    }
  }
  for (int i = tid; i < m.L * 16; i += nthreads) {
    const int l = i >> 4, n = i & 15;
    bias[i] = (m.b_off[l] >= 0 && n < m.dims[l + 1]) ? params[m.b_off[l] + n] : 0.f;
  }
}

// Where the phi-input columns of this lane come from: lane (g, t) holds columns f = 8j + 2t + c (j, c in {0, 1}) of every
// tile, so the decode of `f` is done once per launch.  Code = kind | offset << 2; kind 0: beyond the input (zero), 1: state
// of the destination, 2: state of the source, 3: cached static column.
__device__ __forceinline__ void ode_lane_columns(const OdeArgs& a, int lane, int (&sel)[4]) {
  const int t = lane & 3, dh = a.dx + a.dhs;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int f = 8 * (q >> 1) + 2 * t + (q & 1);
    int kind = 0, off = 0;
    if (f < a.din) {
      if (f < 2 * dh) {
        const int side = f >= dh ? 1 : 0, ff = f - side * dh;
        if (ff < a.dx) { kind = 1 + side; off = ff; } else { kind = 3; off = side * a.dhs + (ff - a.dx); }
      } else {
        kind = 3;
        off = 2 * a.dhs + (f - 2 * dh);
      }
    }
    sel[q] = kind | (off << 2);
  }
}

// phi input of a 16-edge tile in accumulator layout: z[j][i] = column 8j + 2t + (i & 1) of edge e0 + g + 8 (i >> 1), branch
// free (selects + one shared-memory load per value; `u_off` / `stat_off` are float offsets into `smem`).  Rows beyond the
// CTA's edges repeat its last edge: finite values whose results are never stored and whose cotangents are zero.
__device__ __forceinline__ void ode_gather_tile(const OdeArgs& a, const float* __restrict__ smem, int u_off, int stat_off, const OdeEdgeCache& c,
                                                const int (&sel)[4], int e0, int ne, int lane, float (&z)[2][4]) {
  const int g = lane >> 2;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    if (8 * j < a.din) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = min(e0 + g + 8 * (i >> 1), ne - 1), code = sel[2 * j + (i & 1)], kind = code & 3, off = code >> 2;
        const int node = kind == 2 ? c.es[e] : c.et[e];
        const int idx = kind == 3 ? stat_off + e * c.nstat + off : u_off + node * a.dx + off;
        const float v = smem[kind ? idx : 0];
        z[j][i] = kind ? v : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) z[j][i] = 0.f;
    }
  }
}

// one Dense layer on a 16-edge tile: out = act(in W + b), both in accumulator layout
__device__ __forceinline__ void ode_layer_mma(const MlpDev& m, int l, const float4* __restrict__ wf, const float* __restrict__ bias,
                                              int lane, const float (&in)[2][4], float (&out)[2][4]) {
  const int t = lane & 3;
  const int KT = (m.dims[l] + 7) >> 3, NT = (m.dims[l + 1] + 7) >> 3;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const float b0 = bias[l * 16 + 8 * nt + 2 * t], b1 = bias[l * 16 + 8 * nt + 2 * t + 1];
    out[nt][0] = b0; out[nt][1] = b1; out[nt][2] = b0; out[nt][3] = b1;
  }
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    if (ks < KT) {
      uint32_t hi[4], lo[4];
      frag_c_to_a(in[ks], hi, lo);
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
        if (nt < NT) mma_3x(out[nt], hi, lo, wf[((l * 2 + ks) * 2 + nt) * 32 + lane]);
    }
  }
  const int act = m.act[l];
  if (act != NGPDE_ACT_IDENTITY) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) out[nt][i] = ode_act(act, out[nt][i]);
  }
}

// Cluster barrier with release / acquire at cluster scope: what DSMEM stores need.  (cooperative_groups' cluster.sync() puts a
// gpu-scope MEMBAR in front of it.)
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int L>
__global__ void __launch_bounds__(ODE_THREADS, 1) edgeconv_ode_fwd_mma_kernel(const __grid_constant__ OdeArgs a, const OdeMmaSmem Q) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, cta = blockIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = (int)((long long)a.N * cta / a.ncta), n1 = (int)((long long)a.N * (cta + 1) / a.ncta);
  const int k0 = a.rowptr[n0], k1 = a.rowptr[n1], ne = k1 - k0;
  const int dx = a.dx, nown = (n1 - n0) * dx, nd = a.N * dx, max_own = Q.max_own;
  float4* wf = reinterpret_cast<float4*>(sm + Q.off_wf);
  float* bias = sm + Q.off_bias;
  float* msg = sm + Q.off_msg;
  float* ucur = sm + Q.off_ucur;
  float* kst = sm + Q.off_kst;
  float* u0 = sm + Q.off_u0;
  int* rp = reinterpret_cast<int*>(sm + Q.off_rp);
  OdeEdgeCache ec;
  ec.nstat = Q.nstat;
  ec.stat = sm + Q.off_stat;
  ec.es = reinterpret_cast<int*>(sm + Q.off_es);
  ec.et = reinterpret_cast<int*>(sm + Q.off_et);
  ode_fill_cache(a, ec, k0, k1, tid, ODE_THREADS);
  ode_fill_weight_images(a.mlp, a.params, wf, nullptr, bias, tid, ODE_THREADS);
  for (int i = tid; i <= n1 - n0; i += ODE_THREADS) rp[i] = a.rowptr[n0 + i];
  const size_t ubuf = Q.ubuf;
  for (int i = tid; i < nd; i += ODE_THREADS) ucur[i] = a.u[i];
  for (int i = tid; i < nown; i += ODE_THREADS) {
    u0[i] = a.u[n0 * dx + i];
    a.traj[n0 * dx + i] = u0[i];
  }
  // stage-combination coefficients: cf[s][j] = dt a[s+1][j] (dt b[j] for the last stage), j <= s
  __shared__ float cf[ODE_MAXS * ODE_MAXS];
  if (tid < ODE_MAXS * ODE_MAXS) {
    const int ss = tid / ODE_MAXS, j = tid % ODE_MAXS;
    cf[tid] = (ss < a.S && j <= ss) ? a.dt * (ss + 1 == a.S ? a.b[j] : a.a[ss + 1][j]) : 0.f;
  }
  cluster.sync();
  int cur = 0;
  const int n_rt = (ne + 15) >> 4, g = lane >> 2, t = lane & 3;
  int sel[4];
  ode_lane_columns(a, lane, sel);
  ODE_PROF_DECL(8);
  for (int step = 0; step < a.n_steps; ++step) {
    float* tr = a.traj + (size_t)step * a.S * nd;
    for (int s = 0; s < a.S; ++s) {
      const float* uin = ucur + (size_t)cur * ubuf;
      ODE_PROF_START();
      // ---- edges: 16-edge tiles, phi on tensor cores ----
      for (int rt = warp; rt < n_rt; rt += ODE_THREADS / 32) {
        float z[2][4], o[2][4];
        ODE_PROF_START2();
        ode_gather_tile(a, sm, (int)(uin - sm), Q.off_stat, ec, sel, rt * 16, ne, lane, z);
        ODE_PROF2(3);
#pragma unroll
        for (int l = 0; l < L; ++l) {
          ode_layer_mma(a.mlp, l, wf, bias, lane, z, o);
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) z[j][i] = o[j][i];
          ODE_PROF2(4 + (l < 3 ? l : 3));
        }
        // message = columns [0, dx) of the last layer (dx <= 4: tile 0, lanes t < 2)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int e = rt * 16 + g + 8 * (i >> 1), c = 2 * t + (i & 1);
          if (e < ne && c < dx) msg[(size_t)e * dx + c] = z[0][i];
        }
      }
      __syncthreads();
      ODE_PROF(0);
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
#pragma unroll
        for (int j = 0; j < ODE_MAXS; ++j)   // zero coefficients add exactly nothing
          if (j <= s) nxt = fmaf(cf[s * ODE_MAXS + j], kst[(size_t)j * max_own + i], nxt);
        const int gi = n0 * dx + i;
        if (!last) {
          tr[(size_t)(s + 1) * nd + gi] = nxt;
        } else {
          if (step + 1 < a.n_steps) a.traj[(size_t)(step + 1) * a.S * nd + gi] = nxt;
          a.u[gi] = nxt;
          u0[i] = nxt;
        }
        for (int r = 0; r < a.ncta; ++r) cluster.map_shared_rank(unext, r)[gi] = nxt;
      }
      ODE_PROF(1);
      cluster_barrier();
      ODE_PROF(2);
      cur ^= 1;
    }
  }
  ODE_PROF_END(0, 8);
}

__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ int ode_owner(int node, int N, int ncta) {
  int c = (int)(((long long)node * ncta) / N);
  while (c + 1 < ncta && node >= (int)((long long)N * (c + 1) / ncta)) ++c;
  while (c > 0 && node < (int)((long long)N * c / ncta)) --c;
  return c;
}

template <int L>
__global__ void __launch_bounds__(ODE_THREADS, 1) edgeconv_ode_bwd_mma_kernel(const __grid_constant__ OdeArgs a, const OdeMmaSmem Q) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) float sm[];
  const MlpDev& m = a.mlp;
  const int tid = threadIdx.x, cta = blockIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int n0 = (int)((long long)a.N * cta / a.ncta), n1 = (int)((long long)a.N * (cta + 1) / a.ncta);
  const int k0 = a.rowptr[n0], k1 = a.rowptr[n1], ne = k1 - k0;
  const int dx = a.dx, nown = (n1 - n0) * dx, nd = a.N * dx, max_own = Q.max_own;
  const float4* wf = reinterpret_cast<const float4*>(sm + Q.off_wf);
  const float4* wt = reinterpret_cast<const float4*>(sm + Q.off_wt);
  float* bias = sm + Q.off_bias;
  float* ucur = sm + Q.off_ucur;
  int* rp = reinterpret_cast<int*>(sm + Q.off_rp);
  int* tp = reinterpret_cast<int*>(sm + Q.off_tp);
  int* eown = reinterpret_cast<int*>(sm + Q.off_eown);
  float* kb = sm + Q.off_kb;
  float* ub = sm + Q.off_ub;
  float* lam = sm + Q.off_lam;
  float* dte = sm + Q.off_dte;
  float* dsrc = sm + Q.off_dsrc;
  float* zs = sm + Q.off_patch + warp * (2 * 16 * ODE_PS);
  float* gs = zs + 16 * ODE_PS;
  float* acc = sm + Q.off_acc;
  OdeEdgeCache ec;
  ec.nstat = Q.nstat;
  ec.stat = sm + Q.off_stat;
  ec.es = reinterpret_cast<int*>(sm + Q.off_es);
  ec.et = reinterpret_cast<int*>(sm + Q.off_et);
  // ---- prologue ----
  ode_fill_cache(a, ec, k0, k1, tid, ODE_THREADS);
  ode_fill_weight_images(m, a.params, reinterpret_cast<float4*>(sm + Q.off_wf), reinterpret_cast<float4*>(sm + Q.off_wt), bias, tid,
                         ODE_THREADS);
  for (int i = tid; i <= n1 - n0; i += ODE_THREADS) {
    rp[i] = a.rowptr[n0 + i];
    tp[i] = a.tptr[n0 + i];
  }
  for (int i = tid; i < nown; i += ODE_THREADS) lam[i] = a.lam[n0 * dx + i];
  // inverse of the transpose permutation for the edges whose SOURCE this CTA owns: tinv[edge] = transpose position
  int* tinv = reinterpret_cast<int*>(a.desrc);
  {
    const int q0 = a.tptr[n0], q1 = a.tptr[n1];
    for (int q = q0 + tid; q < q1; q += ODE_THREADS) tinv[a.tpos[q]] = q;
  }
  // stage-cotangent coefficients: ca[sn][j] = dt a[j][sn] (j > sn), cb[sn] = dt b[sn]
  __shared__ float ca[ODE_MAXS * ODE_MAXS], cb[ODE_MAXS];
  if (tid < ODE_MAXS * ODE_MAXS) {
    const int sn = tid / ODE_MAXS, j = tid % ODE_MAXS;
    ca[tid] = (j < a.S && j > sn) ? a.dt * a.a[j][sn] : 0.f;
    if (j == 0) cb[sn] = sn < a.S ? a.dt * a.b[sn] : 0.f;
  }
  int idx = a.n_steps * a.S - 1;
  for (int i = tid; i < nd; i += ODE_THREADS) cp_async4(ucur + i, a.traj + (size_t)idx * nd + i);
  cp_async_wait_all();
  __threadfence();
  cluster.sync();
  for (int e = tid; e < ne; e += ODE_THREADS) {
    const int q = __ldcg(tinv + k0 + e);
    const int c = ode_owner(ec.es[e], a.N, a.ncta);
    const int qbase = a.tptr[(int)((long long)a.N * c / a.ncta)];
    eown[e] = (q - qbase) * ODE_MAX_CTAS + c;
  }
  // stage cotangent of the very first stage visited (s = S - 1 of the last step): dt b_s lam (/ deg)
  for (int i = tid; i < nown; i += ODE_THREADS) {
    const int nl = i / dx, deg = rp[nl + 1] - rp[nl];
    float v = a.dt * a.b[a.S - 1] * lam[i];
    if (a.aggr == NGPDE_AGGR_MEAN && deg > 0) v = __fdiv_rn(v, (float)deg);
    kb[i] = v;
  }
  float dw[L][2][4], db[L][2][2];
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) dw[l][j][i] = 0.f;
      db[l][j][0] = db[l][j][1] = 0.f;
    }
  const int n_rt = (ne + 15) >> 4, dh1 = dx + a.dhs;
  int sel[4];
  ode_lane_columns(a, lane, sel);
  int buf = 0, pbuf = 0;
  __syncthreads();
  ODE_PROF_DECL(8);
  for (; idx >= 0; --idx) {
    const int s = idx % a.S;
    const float* uin = ucur + (size_t)buf * Q.ubuf;
    ODE_PROF_START();
    if (idx > 0) {
      float* un = ucur + (size_t)(buf ^ 1) * Q.ubuf;
      for (int i = tid; i < nd; i += ODE_THREADS) cp_async4(un + i, a.traj + (size_t)(idx - 1) * nd + i);
    }
    float* dsw = dsrc + (size_t)pbuf * Q.dsrc_stride;
    for (int rt = warp; rt < n_rt; rt += ODE_THREADS / 32) {
      const int e0 = rt * 16;
      // ---- recompute: z[l] = input of layer l (z[L] = output), accumulator layout ----
      float z[L + 1][2][4];
      ode_gather_tile(a, sm, (int)(uin - sm), Q.off_stat, ec, sel, e0, ne, lane, z[0]);
#pragma unroll
      for (int l = 0; l < L; ++l) ode_layer_mma(m, l, wf, bias, lane, z[l], z[l + 1]);
      // ---- cotangent of the message: the (scaled) stage cotangent of the destination, columns [0, dx) ----
      float gq[2][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int e = e0 + g + 8 * (i >> 1), c = 2 * t + (i & 1);
        float v = 0.f;
        if (e < ne && c < dx) v = kb[(size_t)(ec.et[e] - n0) * dx + c];
        gq[0][i] = v;
        gq[1][i] = 0.f;
      }
#pragma unroll
      for (int l = L - 1; l >= 0; --l) {
        const int act = m.act[l];
        const int KT = (m.dims[l] + 7) >> 3, NT = (m.dims[l + 1] + 7) >> 3;
        if (act != NGPDE_ACT_IDENTITY) {
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) gq[j][i] *= act_grad_y(act, z[l + 1][j][i]);
        }
        // bias gradient: this lane's two edges, columns 2t / 2t + 1 of every tile
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          db[l][j][0] += gq[j][0] + gq[j][2];
          db[l][j][1] += gq[j][1] + gq[j][3];
        }
        // weight gradient dW_l += Z_l' G_l: transpose both through the warp's patch ([feature][edge])
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int f = 8 * j + 2 * t + (i & 1), e = g + 8 * (i >> 1);
            zs[f * ODE_PS + e] = z[l][j][i];
            gs[f * ODE_PS + e] = gq[j][i];
          }
        __syncwarp();
        // the tensor-core accumulator truncates: a launch-long running sum in it drifts (3e-5 after 1,440 MMAs); the tile's
        // 16-edge product starts from zero and joins the running sum by a round-to-nearest FP32 add
        float dwt[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) dwt[nt][i] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t hi[4], lo[4];
          split_tf32(zs[g * ODE_PS + 8 * ks + t], hi[0], lo[0]);
          split_tf32(zs[(g + 8) * ODE_PS + 8 * ks + t], hi[1], lo[1]);
          split_tf32(zs[g * ODE_PS + 8 * ks + t + 4], hi[2], lo[2]);
          split_tf32(zs[(g + 8) * ODE_PS + 8 * ks + t + 4], hi[3], lo[3]);
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            if (nt < NT) {
              uint32_t b0h, b0l, b1h, b1l;
              split_tf32(gs[(8 * nt + g) * ODE_PS + 8 * ks + t], b0h, b0l);
              split_tf32(gs[(8 * nt + g) * ODE_PS + 8 * ks + t + 4], b1h, b1l);
              mma_tf32(dwt[nt], lo, b0h, b1h);
              mma_tf32(dwt[nt], hi, b0l, b1l);
              mma_tf32(dwt[nt], hi, b0h, b1h);
            }
          }
        }
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int i = 0; i < 4; ++i) dw[l][nt][i] += dwt[nt][i];
        // input gradient dH = G W'
        float dh[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) dh[j][i] = 0.f;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (j < NT) {
            uint32_t hi[4], lo[4];
            frag_c_to_a(gq[j], hi, lo);
#pragma unroll
            for (int kt = 0; kt < 2; ++kt)
              if (kt < KT) mma_3x(dh[kt], hi, lo, wt[((l * 2 + j) * 2 + kt) * 32 + lane]);
          }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) gq[j][i] = dh[j][i];
      }
      // gq = d/d(phi input): columns [0, dx) -> destination side (kept here), [dh1, dh1 + dx) -> the source's owner
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int e = e0 + g + 8 * (i >> 1), f = 8 * j + 2 * t + (i & 1);
          if (e < ne) {
            if (f < dx) dte[(size_t)e * dx + f] = gq[j][i];
            if (f >= dh1 && f < dh1 + dx) {
              const int pk = eown[e];
              cluster.map_shared_rank(dsw, pk % ODE_MAX_CTAS)[(size_t)(pk / ODE_MAX_CTAS) * dx + (f - dh1)] = gq[j][i];
            }
          }
        }
    }
    cp_async_wait_all();
    ODE_PROF(1);
    cluster_barrier();  // every CTA's source-side cotangents of this stage have landed; the next stage input is in place
    ODE_PROF(3);
    // ---- owned nodes: ubar_s = destination side (CSR order) + source side (transpose order); next stage cotangent ----
    for (int i = tid; i < nown; i += ODE_THREADS) {
      const int nl = i / dx, c = i - nl * dx;
      float v = 0.f;
      for (int k = rp[nl] - k0; k < rp[nl + 1] - k0; ++k) v += dte[(size_t)k * dx + c];
      const int q0 = tp[0];
      for (int q = tp[nl] - q0; q < tp[nl + 1] - q0; ++q) v += dsw[(size_t)q * dx + c];
      ub[(size_t)s * max_own + i] = v;
      int sn = s - 1;
      if (s == 0) {  // step finished: lam <- lam + sum_s ubar_s
        float l2 = lam[i];
        for (int j = 0; j < a.S; ++j) l2 += ub[(size_t)j * max_own + i];
        lam[i] = l2;
        sn = a.S - 1;
      }
      // kbar_sn = dt b_sn lam + dt sum_{j > sn} a_j,sn ubar_j   (the ubar of the step being entered: none yet when sn = S - 1)
      float kv = cb[sn] * lam[i];
#pragma unroll
      for (int j = 1; j < ODE_MAXS; ++j)   // zero coefficients (and stages that do not exist) add exactly nothing
        if (j > sn && j < a.S) kv = fmaf(ca[sn * ODE_MAXS + j], ub[(size_t)j * max_own + i], kv);
      const int deg = rp[nl + 1] - rp[nl];
      if (a.aggr == NGPDE_AGGR_MEAN && deg > 0) kv = __fdiv_rn(kv, (float)deg);
      kb[i] = kv;
    }
    __syncthreads();
    ODE_PROF(4);
    buf ^= 1;
    pbuf ^= 1;
  }
  ODE_PROF_END(8, 6);
  for (int i = tid; i < nown; i += ODE_THREADS) a.lam[n0 * dx + i] = lam[i];
  // ---- parameter gradient: lanes of equal t hold the bias partials of different edges -> fixed butterfly; warps are added
  // in warp order into the CTA's accumulator, CTAs in CTA order by CTA 0 ----
#pragma unroll
  for (int l = 0; l < L; ++l)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v = db[l][j][c];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        db[l][j][c] = v;
      }
  for (int p = tid; p < m.n_params; p += ODE_THREADS) acc[p] = 0.f;
  __syncthreads();
  for (int w = 0; w < ODE_THREADS / 32; ++w) {
    if (warp == w) {
#pragma unroll
      for (int l = 0; l < L; ++l) {
        const int K = m.dims[l], N = m.dims[l + 1];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k = g + 8 * (i >> 1), n = 8 * j + 2 * t + (i & 1);
            if (k < K && n < N) acc[m.w_off[l] + k * N + n] += dw[l][j][i];
          }
          if (g == 0 && m.b_off[l] >= 0) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int n = 8 * j + 2 * t + c;
              if (n < N) acc[m.b_off[l] + n] += db[l][j][c];
            }
          }
        }
      }
    }
    __syncthreads();
  }
  for (int p = tid; p < m.n_params; p += ODE_THREADS) a.dpart[(size_t)cta * m.n_params + p] = acc[p];
  __threadfence();
  cluster.sync();
  if (cta == 0) {
    for (int p = tid; p < m.n_params; p += ODE_THREADS) {
      float s = 0.f;
      for (int c = 0; c < a.ncta; ++c) s += __ldcg(a.dpart + (size_t)c * m.n_params + p);
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
  P->off_dpart = off; off = (off + sizeof(float) * ODE_MAX_CTAS * a.mlp.n_params + 255) & ~size_t(255);
  P->off_wpad = off;  off = (off + sizeof(c_ode_w) + 255) & ~size_t(255);
  P->total = off + 256;
  return NGPDE_OK;
}

// largest per-CTA edge count (and transposed-edge count) for a cluster of `ncta`: the row pointers live on the device -> read
// the boundary entries once per graph
int ode_max_edges(ngpde_graph* g, cudaStream_t st, int ncta, int* out, int* out_t = nullptr) {
  const int slot = ncta == ODE_MAX_CTAS ? 1 : 0;
  if (g->ode_max_edges[slot] < 0) {
    int h[2][ODE_MAX_CTAS + 1];
    for (int c = 0; c <= ncta; ++c) {
      const int n = (int)((long long)g->N * c / ncta);
      NGPDE_CUDA_TRY(cudaMemcpyAsync(&h[0][c], g->rowptr + n, sizeof(int), cudaMemcpyDeviceToHost, st));
      NGPDE_CUDA_TRY(cudaMemcpyAsync(&h[1][c], g->tptr + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    NGPDE_CUDA_TRY(cudaStreamSynchronize(st));
    int m = 0, mt = 0;
    for (int c = 0; c < ncta; ++c) {
      m = std::max(m, h[0][c + 1] - h[0][c]);
      mt = std::max(mt, h[1][c + 1] - h[1][c]);
    }
    g->ode_max_edges[slot] = m;
    g->ode_max_tedges[slot] = mt;
  }
  *out = g->ode_max_edges[slot];
  if (out_t) *out_t = g->ode_max_tedges[slot];
  return NGPDE_OK;
}

// padded parameters -> constant bank (device-to-device, in stream order)
int ode_stage_params(const OdeArgs& a, int W, float* staging, cudaStream_t st) {
  ode_pad_params_kernel<<<8, 256, 0, st>>>(a.mlp, a.params, W, staging);
  NGPDE_CUDA_TRY(cudaGetLastError());
  NGPDE_CUDA_TRY(cudaMemcpyToSymbolAsync(c_ode_w, staging, sizeof(float) * (size_t)a.mlp.L * (W * W + W), 0, cudaMemcpyDeviceToDevice, st));
  return NGPDE_OK;
}

template <class K>
void cluster_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, K kernel, int ncta, size_t smem, cudaStream_t st) {
  *cfg = cudaLaunchConfig_t{};
  cfg->gridDim = dim3(ncta);
  cfg->blockDim = dim3(ODE_THREADS);
  cfg->dynamicSmemBytes = smem;
  cfg->stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = ncta;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg->attrs = attr;
  cfg->numAttrs = 1;
}

template <class K, class... Args>
int launch_cluster(K kernel, int ncta, size_t smem, cudaStream_t st, Args... args) {
  NGPDE_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (ncta > 8) NGPDE_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  cluster_config(&cfg, attr, kernel, ncta, smem, st);
  NGPDE_CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, args...));
  return NGPDE_OK;
}

// can a cluster of `ncta` CTAs with this much shared memory be co-scheduled at all?  (16 is a non-portable size: it needs a
// GPC with 16 free SMs; MIG slices and some floor-swept parts do not have one)
template <class K>
bool cluster_fits(K kernel, int ncta, size_t smem) {
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return false; }
  if (ncta > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return false; }
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  cluster_config(&cfg, attr, kernel, ncta, smem, nullptr);
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
  return n >= 1;
}

template <template <int> class F, class... Args>
int ode_dispatch_layers(int L, Args... args) {
  switch (L) {
    case 1: return F<1>::run(args...);
    case 2: return F<2>::run(args...);
    case 3: return F<3>::run(args...);
    default: return F<4>::run(args...);
  }
}
template <int L> struct OdeFwdMma {
  static int run(bool probe, size_t smem, cudaStream_t st, const OdeArgs& a, const OdeMmaSmem& q) {
    if (probe) return cluster_fits(edgeconv_ode_fwd_mma_kernel<L>, a.ncta, smem) ? 1 : 0;
    return launch_cluster(edgeconv_ode_fwd_mma_kernel<L>, a.ncta, smem, st, a, q);
  }
};
template <int L> struct OdeBwdMma {
  static int run(bool probe, size_t smem, cudaStream_t st, const OdeArgs& a, const OdeMmaSmem& q) {
    if (probe) return cluster_fits(edgeconv_ode_bwd_mma_kernel<L>, a.ncta, smem) ? 1 : 0;
    return launch_cluster(edgeconv_ode_bwd_mma_kernel<L>, a.ncta, smem, st, a, q);
  }
};

// the tensor-core kernels on the widest cluster the device grants: 16 CTAs (non-portable size) when the graph has the nodes
// for it and such a cluster can be scheduled, else the portable 8
template <template <int> class F>
int ode_launch_mma(ngpde_graph* g, OdeArgs& a, bool adjoint, cudaStream_t st) {
  for (int ncta : {ODE_MAX_CTAS, ODE_CTAS}) {
    if (ncta == ODE_MAX_CTAS && (a.N < 256 || g->ode_wide_cluster[adjoint] == 0)) continue;
    a.ncta = ncta;
    int max_tedges = 0;
    if (int rc = ode_max_edges(g, st, ncta, &a.max_edges, &max_tedges)) return rc;
    const OdeMmaSmem q = ode_mma_smem(a, adjoint, max_tedges);
    const size_t smem = sizeof(float) * (size_t)q.floats;
    if (ncta == ODE_MAX_CTAS) {
      if (g->ode_wide_cluster[adjoint] < 0 || smem > 220 * 1024) {
        const bool ok = smem <= 220 * 1024 && ode_dispatch_layers<F>(a.mlp.L, true, smem, st, a, q) == 1;
        if (smem <= 220 * 1024) g->ode_wide_cluster[adjoint] = ok ? 1 : 0;
        if (!ok) continue;
      }
    }
    NGPDE_REQUIRE(smem <= 220 * 1024, "persistent ODE kernel: the graph (%d nodes, %d edges per CTA) does not fit shared memory; use the "
                  "CUDA-graph step path", a.N, a.max_edges);
    return ode_dispatch_layers<F>(a.mlp.L, false, smem, st, a, q);
  }
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
  char* ws = static_cast<char*>(workspace);
  a.params = phi_params; a.snode = snode; a.dt = dt; a.n_steps = n_steps; a.u = u; a.traj = traj;
  a.kbuf = reinterpret_cast<float*>(ws + P.off_kbuf);
  a.dbg = tc_get_debug_buffer();
  const int W = ode_width(a.mlp);
  if (W == 16) return ode_launch_mma<OdeFwdMma>(g, a, false, st);  // tensor-core kernels
  a.ncta = ODE_CTAS;
  if (int rc = ode_max_edges(g, st, a.ncta, &a.max_edges)) return rc;
  const int nd = a.N * a.dx, max_own = ((a.N + ODE_CTAS - 1) / ODE_CTAS + 1) * a.dx, nstat = 2 * a.dhs + a.dpos;
  const size_t fl = (((size_t)a.max_edges * a.dx + 3) & ~size_t(3)) + 2 * (((size_t)nd + 3) & ~size_t(3)) +
                    (size_t)a.S * max_own + max_own + (((size_t)a.max_edges * nstat + 3) & ~size_t(3)) + 2 * (size_t)a.max_edges +
                    (size_t)max_own + 8;
  const size_t smem = sizeof(float) * fl;
  NGPDE_REQUIRE(smem <= 200 * 1024, "ode_forward: the graph (%d nodes, %d edges per CTA) does not fit shared memory; use the "
                "CUDA-graph step path", a.N, a.max_edges);
  if (int rc = ode_stage_params(a, W, reinterpret_cast<float*>(ws + P.off_wpad), st)) return rc;
  return launch_cluster(edgeconv_ode_fwd_kernel<32>, a.ncta, smem, st, a);
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
  char* ws = static_cast<char*>(workspace);
  a.params = phi_params; a.snode = snode; a.dt = dt; a.n_steps = n_steps;
  a.traj = const_cast<float*>(traj); a.lam = lam; a.dparams = dphi_params;
  a.kbuf = reinterpret_cast<float*>(ws + P.off_kbuf);
  a.desrc = reinterpret_cast<float*>(ws + P.off_desrc);
  a.dpart = reinterpret_cast<float*>(ws + P.off_dpart);
  a.dbg = tc_get_debug_buffer();
  const int W = ode_width(a.mlp);
  if (W == 16) return ode_launch_mma<OdeBwdMma>(g, a, true, st);  // tensor-core kernels
  a.ncta = ODE_CTAS;
  if (int rc = ode_max_edges(g, st, a.ncta, &a.max_edges)) return rc;
  const OdeBwdSmem L = ode_bwd_smem(a.mlp, W, a.dx, P.max_nodes, a.max_edges, 2 * a.dhs + a.dpos);
  const size_t smem = sizeof(float) * (size_t)L.floats + 64;
  NGPDE_REQUIRE(smem <= 220 * 1024, "ode_adjoint: the tile buffers need %zu bytes of shared memory", smem);
  if (int rc = ode_stage_params(a, W, reinterpret_cast<float*>(ws + P.off_wpad), st)) return rc;
  return launch_cluster(edgeconv_ode_bwd_kernel<32>, a.ncta, smem, st, a, L);
}
