// Tensor-core (tcgen05 + TMEM) message-passing kernels: eligibility, layouts and launches.  Kept in its own translation
// unit so that the FP32-FFMA engine (ngpde_conv.cu) and these kernels compile independently.
#include <algorithm>
#include <cstring>
#include <cstdlib>

#include "ngpde_tc.cuh"
#include "ngpde_tc_bwd.cuh"

namespace ngpde {
namespace {

// ---- tensor-core (tcgen05) path selection ----
constexpr int kSmemMaxTc = 227 * 1024;
bool g_use_tc = true;
long long* g_tcb_dbg = nullptr;  // NGPDE_OPT_DEBUG_BUFFER: phase timestamps of the tensor-core backward (edge phase)

int pad16(int x) { return (x + 15) / 16 * 16; }

// Shapes and prepared-weight-block offsets of `m` on the tensor-core path; false when a layer is too wide for it.
bool tc_fill_layout(const MlpDev& m, TcLayout* lay) {
  if (m.L < 1 || m.L > NGPDE_MAX_LAYERS) return false;
  *lay = TcLayout{};
  lay->L = m.L;
  int off = 0, kmax = 0;
  for (int l = 0; l < m.L; ++l) {
    const int K = m.dims[l], N = m.dims[l + 1];
    if (N > TC_MAXN || K > 192) return false;
    lay->K[l] = K;
    lay->N[l] = N;
    lay->Kd[l] = pad16(K);
    lay->Kp[l] = lay->Kd[l] + 8;
    lay->Np[l] = pad16(N);
    lay->img_floats[l] = ((lay->Np[l] + 31) / 32) * lay->Kp[l] * 32;
    lay->img_off[l] = off;
    off += 2 * lay->img_floats[l];  // Kp is a multiple of 8, so every image is a multiple of 1 KB
    kmax = std::max(kmax, lay->Kp[l]);
  }
  lay->block_floats = off;
  lay->kmax = kmax;
  lay->cols_group = TC_MAXN + 2 * kmax;
  return true;
}

bool tc_bwd_make_impl(const MlpDev& m, bool contract, bool addend, bool node, int aggr, bool need_dz0, TcBwdPhase* t) {
  *t = TcBwdPhase{};
  if (!g_use_tc || contract || addend || m.L > TCB_MAXL) return false;
  if (!node && !(aggr == NGPDE_AGGR_SUM || aggr == NGPDE_AGGR_MEAN)) return false;
  if (!tc_fill_layout(m, &t->lay)) return false;
  const TcLayout& lay = t->lay;
  const int L = lay.L;
  if (m.act[L - 1] != NGPDE_ACT_IDENTITY) return false;  // Z_L is not recomputed
  for (int l = 0; l < L; ++l) {
    if (!act_grad_from_y(m.act[l])) return false;         // swish / gelu need the pre-activation
    if (lay.Kd[l] > 80) return false;                     // register accumulators cover 64 + 16 columns of dW^T
  }
  // TMEM columns: FP32 copies of Z_1..Z_{L-1} | A hi, A lo | D | Dw.  D / Dw double as the two accumulators of a
  // recomputed layer (two issuing warps), so they are as wide as the widest hidden output as well.
  int zs = 0, kdmax = 0, awidth = 0, dmax = 0, wmax = 0;
  for (int l = 1; l < L; ++l) {
    t->c_zs[l] = zs;
    zs += lay.Kd[l];
    dmax = std::max(dmax, lay.Kd[l]);   // dZ_l
    wmax = std::max(wmax, lay.Kd[l]);   // dW_l^T has Kd_l columns
  }
  for (int l = 0; l < L; ++l) {
    kdmax = std::max(kdmax, lay.Kd[l]);
    awidth = std::max(awidth, std::max(lay.Kd[l], lay.Np[l]));
    if (l < L - 1) {
      dmax = std::max(dmax, lay.Np[l]);
      wmax = std::max(wmax, lay.Np[l]);
    }
  }
  const bool alias0 = zs >= 2 * lay.Kd[0];  // layer 0's outputs may overwrite the (dead by then) activation copies
  if (!alias0) {
    dmax = std::max(dmax, lay.Kd[0]);
    wmax = std::max(wmax, lay.Kd[0]);
  }
  t->a_width = awidth;
  t->c_a = zs;
  t->c_d = zs + 2 * awidth;
  t->c_dw = t->c_d + dmax;
  t->c_d0 = alias0 ? 0 : t->c_d;
  t->c_dw0 = alias0 ? lay.Kd[0] : t->c_dw;
  int total = t->c_dw + wmax;
  if (total > 512) return false;
  if (total + wmax <= 512 && L > 1) {  // room for a second dW^T accumulator: collection moves off the critical path
    t->dw_alt = wmax;
    total += wmax;
  }
  int cols = 32;
  while (cols < total) cols *= 2;
  t->tmem_cols = cols;
  // shared memory: weight images | column tables | staging (Z hi, Z lo: nzh groups each; G hi, G lo: 2 groups each) | dZ_0.
  // Preferred: stage the weight-gradient operands for the whole 128-row tile at once (all 16 warps work, one MMA batch
  // per layer).  That needs (2 nzh + 4) x 16 KB of staging; when the weight images do not fit beside it, the images of
  // layer 1 (used by the recompute and, late, by its own input gradient) and of the last layer (used once, first thing in
  // the backward sweep) share one slot and are swapped by TMA twice per tile.  Otherwise: two 64-row halves, all resident.
  t->nzh = (kdmax + 31) / 32;
  t->nzl = t->nzh;
  const int dz_bytes = (!node && need_dz0) ? TC_TILE * (lay.Kd[0] + 1) * 4 : 0;
  const int cols_bytes = (int)(sizeof(TcCol) + sizeof(TcDst)) * lay.Kd[0] + 2 * TC_MAX_CHUNKS * (int)sizeof(TcChunk);
  auto place = [&](bool full, bool stream) -> bool {
    int off = 0;  // floats
    const int sa = stream ? 1 : -1, sb = stream ? L - 1 : -1;
    int slot = 0;
    for (int l = 0; l < L; ++l) {
      if (l == sb) {
        t->woff[l] = t->woff[sa];
        continue;
      }
      t->woff[l] = off;
      int fl = 2 * lay.img_floats[l];
      if (l == sa) fl = slot = std::max(fl, 2 * lay.img_floats[sb]);
      off += fl;
    }
    (void)slot;
    t->stream_a = sa;
    t->stream_b = sb;
    t->full = full;
    t->off_cols = 4 * off;
    t->off_stage = (t->off_cols + cols_bytes + 1023) & ~1023;
    t->off_dz = t->off_stage + (2 * t->nzh + 4) * (full ? TC_TILE : TCB_HALF) * 128;
    const int totalb = 1024 + t->off_dz + dz_bytes;
    t->smem = std::max(totalb, 116 * 1024);  // one CTA per SM: its TMEM allocation must not wait for a neighbour's
    return totalb <= kSmemMaxTc;
  };
  const char* force = getenv("NGPDE_TCB_STAGING");  // developer switch: "half" forces the two-pass staging
  const bool allow_full = !(force && force[0] == 'h') && kdmax <= 64;
  bool ok = allow_full && place(true, false);
  if (!ok && allow_full && L >= 4) ok = place(true, true);
  if (!ok) ok = place(false, false);
  if (!ok) return false;
  t->on = true;
  return true;
}

// Lays out the prepared weight block of `m` and decides whether the tcgen05 forward kernels can run it.
bool tc_make_layout_impl(const MlpDev& m, bool contract, bool addend, int dout, bool node, TcLayout* lay, int* smem_bytes,
                    int* off_cols, int* off_groups, int* group_bytes) {
  if (!g_use_tc || contract || addend || m.L < 1 || m.L > NGPDE_MAX_LAYERS) return false;
  if (!tc_fill_layout(m, lay)) return false;
  const int need = TC_GROUPS * lay->cols_group;
  if (need > 512) return false;
  int cols = 32;
  while (cols < need) cols *= 2;
  lay->tmem_cols = cols;
  *off_cols = 4 * lay->block_floats;
  *off_groups = (*off_cols + (int)sizeof(TcCol) * lay->Kd[0] + TC_MAX_CHUNKS * (int)sizeof(TcChunk) + 127) & ~127;
  *group_bytes = node ? 128 : ((TC_TILE * (dout + 1) * 4 + 127) & ~127);
  const int total = 1024 + *off_groups + TC_GROUPS * *group_bytes;
  // at least half of the SM's shared memory, so that two CTAs (and two 512-column TMEM allocations) never share an SM
  *smem_bytes = std::max(total, 116 * 1024);
  return total <= kSmemMaxTc;
}


// Node phase: put the input segments whose width is a multiple of 16 columns first (stable), so that a 16-column chunk of the
// on-chip layer-0 input is an aligned run of one array (vector loads / stores, see TcChunk).  The weight image rows and the
// parameter-gradient rows of layer 0 follow the same order through TcLayout::ps/po/pw.
void tc_permute_node_segs(Seg* segs, int n_segs, TcLayout* lay) {
  lay->np = 0;
  if (n_segs <= 1 || n_segs > 8) return;
  for (int i = 0; i < n_segs; ++i)
    for (int j = i + 1; j < n_segs; ++j)
      if (segs[i].row < segs[j].row + segs[j].width && segs[j].row < segs[i].row + segs[i].width) return;  // summed segments stay put
  Seg out[8];
  int n = 0, row = 0;
  for (int pass = 0; pass < 2; ++pass)
    for (int i = 0; i < n_segs; ++i)
      if (((segs[i].width & 15) == 0) == (pass == 0)) {
        out[n] = segs[i];
        lay->ps[n] = row;
        lay->po[n] = segs[i].row;
        lay->pw[n] = segs[i].width;
        out[n].row = row;
        row += segs[i].width;
        ++n;
      }
  bool identity = true;
  for (int i = 0; i < n; ++i) identity = identity && lay->ps[i] == lay->po[i];
  if (identity) return;
  lay->np = n;
  for (int i = 0; i < n; ++i) segs[i] = out[i];
}

template <bool NODE>
int launch_fwd_tc_t(int num_sms, const TcPhase& t, const MlpDev& mlp, const float* params, const FwdArgs& base,
                  float* wblock, cudaStream_t st) {
  if (base.tg.n_units <= 0) return NGPDE_OK;
  TcFwdArgs a{};
  a.tg = base.tg;
  std::memcpy(a.arr, base.arr, sizeof(a.arr));
  std::memcpy(a.ld, base.ld, sizeof(a.ld));
  a.n_segs = base.n_segs;
  std::memcpy(a.segs, base.segs, sizeof(a.segs));
  a.lay = t.lay;
  if (NODE) tc_permute_node_segs(a.segs, a.n_segs, &a.lay);
  tc_prep_weights_kernel<<<32, 256, 0, st>>>(params, mlp, a.lay, wblock);
  for (int l = 0; l < mlp.L; ++l) a.act[l] = mlp.act[l];
  a.wblock = wblock;
  a.aggr = base.aggr;
  a.dout = base.dout;
  a.out = base.out;
  a.out_ld = base.out_ld > 0 ? base.out_ld : base.dout;
  a.skip_l0 = base.skip_l0;
  a.off_cols = t.off_cols;
  a.off_groups = t.off_groups;
  a.group_bytes = t.group_bytes;
  NGPDE_CUDA_TRY(cudaFuncSetAttribute(mp_fwd_tc_kernel<NODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, t.smem));
  const int grid = std::max(1, std::min((a.tg.n_units + TC_GROUPS - 1) / TC_GROUPS, num_sms));
  mp_fwd_tc_kernel<NODE><<<grid, TC_THREADS, t.smem, st>>>(a);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

template <bool NODE>
int launch_bwd_tc_t(const TcBwdPhase& t, const MlpDev& mlp, const BwdArgs& base, float* wblock, cudaStream_t st) {
  if (base.tg.n_units <= 0) return NGPDE_OK;
  TcBwdArgs a{};
  a.tg = base.tg;
  std::memcpy(a.arr, base.arr, sizeof(a.arr));
  std::memcpy(a.ld, base.ld, sizeof(a.ld));
  a.n_segs = base.n_segs;
  std::memcpy(a.segs, base.segs, sizeof(a.segs));
  a.lay = t.lay;
  if (NODE) tc_permute_node_segs(a.segs, a.n_segs, &a.lay);
  tc_prep_weights_kernel<<<32, 256, 0, st>>>(base.params, mlp, a.lay, wblock);
  for (int l = 0; l < mlp.L; ++l) {
    a.act[l] = mlp.act[l];
    a.w_off[l] = mlp.w_off[l];
    a.b_off[l] = mlp.b_off[l];
  }
  a.n_params = mlp.n_params;
  a.wblock = wblock;
  a.aggr = base.aggr;
  a.dout = base.dout;
  a.gout_ptr = base.gout_ptr;
  a.gout_ld = base.gout_ld > 0 ? base.gout_ld : base.dout;
  a.skip_w0 = base.skip_w0;
  a.yact = base.yact;
  a.yact_kind = base.yact_kind;
  a.direct_src = base.direct_src;
  a.src_c0 = base.src_w > 0 ? base.src_c0 : 0;
  a.src_w = base.src_w > 0 ? base.src_w : base.dx;
  a.dst_c0 = base.dst_w > 0 ? base.dst_c0 : 0;
  a.dst_w = base.dst_w > 0 ? base.dst_w : base.dx;
  a.dparams_partial = base.dparams_partial;
  a.dx_direct = base.dx_direct;
  a.dmbar = base.dmbar;
  a.dxdst = base.dxdst;
  a.desrc = base.desrc;
  a.dx = base.dx;
  a.need_dz0 = base.need_dz0;
  a.has_dst_side = base.has_dst_side;
  std::memcpy(a.c_zs, t.c_zs, sizeof(a.c_zs));
  a.c_a = t.c_a; a.a_width = t.a_width; a.c_d = t.c_d; a.c_dw = t.c_dw; a.c_d0 = t.c_d0; a.c_dw0 = t.c_dw0; a.tmem_cols = t.tmem_cols; a.dw_alt = t.dw_alt;
  a.off_cols = t.off_cols; a.off_stage = t.off_stage; a.off_dz = t.off_dz; a.nzh = t.nzh; a.nzl = t.nzl;
  std::memcpy(a.woff, t.woff, sizeof(a.woff));
  a.stream_a = t.stream_a; a.stream_b = t.stream_b;
  a.dbg = (NODE != (getenv("NGPDE_TCB_DBG_NODE") != nullptr)) ? nullptr : g_tcb_dbg;
  { const char* e = getenv("NGPDE_TCB_OPT"); a.opt = e ? atoi(e) : 0; }
  // the extended variant only where its extras are needed (hoisted first layers, wide cotangent rows, GCNConv's fused act')
  const bool ext = a.skip_w0 || a.direct_src || a.yact != nullptr || (!NODE && a.need_dz0 && (a.src_w >= 32 || a.dst_w >= 32));
  auto go = [&](auto kernel) -> int {
    NGPDE_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, t.smem));
    kernel<<<t.grid, TCB_THREADS, t.smem, st>>>(a);
    return NGPDE_OK;
  };
  int rc;
  if (t.full) rc = ext ? go(mp_bwd_tc_kernel<NODE, true, true>) : go(mp_bwd_tc_kernel<NODE, true, false>);
  else rc = ext ? go(mp_bwd_tc_kernel<NODE, false, true>) : go(mp_bwd_tc_kernel<NODE, false, false>);
  if (rc) return rc;
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

}  // namespace

bool tc_bwd_make(const MlpDev& m, bool contract, bool addend, bool node, int aggr, bool need_dz0, TcBwdPhase* t) {
  return tc_bwd_make_impl(m, contract, addend, node, aggr, need_dz0, t);
}
bool tc_make_layout(const MlpDev& m, bool contract, bool addend, int dout, bool node, TcLayout* lay, int* smem_bytes,
                    int* off_cols, int* off_groups, int* group_bytes) {
  return tc_make_layout_impl(m, contract, addend, dout, node, lay, smem_bytes, off_cols, off_groups, group_bytes);
}
int launch_fwd_tc(bool node, int num_sms, const TcPhase& t, const MlpDev& mlp, const float* params, const FwdArgs& base,
                  float* wblock, cudaStream_t st) {
  return node ? launch_fwd_tc_t<true>(num_sms, t, mlp, params, base, wblock, st)
              : launch_fwd_tc_t<false>(num_sms, t, mlp, params, base, wblock, st);
}
int launch_bwd_tc(bool node, const TcBwdPhase& t, const MlpDev& mlp, const BwdArgs& base, float* wblock, cudaStream_t st) {
  return node ? launch_bwd_tc_t<true>(t, mlp, base, wblock, st) : launch_bwd_tc_t<false>(t, mlp, base, wblock, st);
}
void tc_set_enabled(bool on) { g_use_tc = on; }
bool tc_get_enabled() { return g_use_tc; }
void tc_set_debug_buffer(long long* p) { g_tcb_dbg = p; }
long long* tc_get_debug_buffer() { return g_tcb_dbg; }

}  // namespace ngpde
// 
