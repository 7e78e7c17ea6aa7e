// Tensor-core (tcgen05 + TMEM) variant of the fused message-passing kernels for MLPs whose layers are at most 64 wide.
//
// One thread owns one edge (edge phase) or one node (node phase): a group of 128 threads owns a tile of 128 rows, which
// is the M dimension of every MMA.  Activations never leave the SM and never touch shared memory:
//     gather -> registers -> TMEM (A operand, lane = row, column = feature, split hi/lo for 3xTF32)
//     tcgen05.mma  D[128 x N](TMEM) = A(TMEM) x W(smem image)           3 passes: lo*hi + hi*lo + hi*hi, FP32 accumulate
//     tcgen05.ld D -> registers: + bias, activation (accurate tanhf/expf), split -> tcgen05.st -> next layer's A
// The last layer's rows go through a padded shared-memory tile only to be reduced per destination node in stored edge
// order (same atomic-free sequential reduction as the FFMA engine), or straight to global memory in the node phase.
// Weights are prepared once per call as SWIZZLE_128B_BASE32B images (tc_prep_weights_kernel) and pulled into shared
// memory with one TMA bulk copy per CTA; two 128-thread groups per CTA keep two tiles in flight so one group's
// epilogue overlaps the other's MMAs.  Operand conventions are pinned by tools/umma_probe.cu.
#pragma once
#include "ngpde_conv.cuh"
#include "ngpde_tc_layout.cuh"
#include "ngpde_umma.cuh"

namespace ngpde {

// how column `c` of the MLP input is produced from the gathered arrays
struct TcCol {
  const float* base;  // array + column offset
  int ld;
  int kind;           // SEG_*
};

struct TcFwdArgs {
  TileGraph tg;
  const float* arr[ARR_COUNT];
  int ld[ARR_COUNT];
  int n_segs;
  Seg segs[8];
  TcLayout lay;
  int act[NGPDE_MAX_LAYERS];
  const float* wblock;  // prepared weight block in global memory (tc_prep_weights_kernel)
  int aggr;
  int dout;
  float* out;           // edge phase: mbar [N][dout]; node phase: y [N][out_ld]
  int out_ld;
  int skip_l0;          // layer 0 = identity (hoisted first layer): activation at the gather, the layer loop starts at 1
  int off_cols;         // byte offsets into dynamic shared memory: column table, per-group regions
  int off_groups;
  int group_bytes;
};

// ---- prepared weight block: per layer hi image then lo image (SWIZZLE_128B_BASE32B, rows = k, bias at row Kd) ----
__global__ void tc_prep_weights_kernel(const float* __restrict__ params, MlpDev mlp, TcLayout lay, float* __restrict__ out) {
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int l = 0; l < lay.L; ++l) {
    const int K = lay.K[l], N = lay.N[l], Kp = lay.Kp[l], Kd = lay.Kd[l];
    const int cols = lay.img_floats[l] / Kp;  // 32 * groups
    const float* W = params + mlp.w_off[l];
    const float* b = mlp.b_off[l] >= 0 ? params + mlp.b_off[l] : nullptr;
    float* hi = out + lay.img_off[l];
    float* lo = hi + lay.img_floats[l];
    for (int i = t0; i < Kp * cols; i += stride) {
      const int k = i / cols, n = i - k * cols;
      float w = 0.f;
      if (n < N) {
        if (k < K) w = W[(size_t)(l == 0 ? tc_orig_row(lay, k) : k) * N + n];
        else if (k == Kd && b != nullptr) w = b[n];
      }
      const float h = umma::tf32_hi(w);
      const uint32_t off = umma::sw128b32_offset(n >> 5, Kp, k, n & 31);
      hi[off] = h;
      lo[off] = umma::tf32_hi(w - h);
    }
  }
}

__device__ __forceinline__ void group_bar(int grp) {
  asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(TC_GTHREADS) : "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   umma::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(umma::smem_u32(bar))
               : "memory");
}

// column table of the gather: cols[c] for c < Kd[0] (padding columns get kind = -1).
// Two segments may cover the SAME input rows (the hoisted first layer, ngpde_conv.cu: a destination-side and a source-side
// projection of one array): the column is then their sum, kind = TC_KIND_DPS with the float distance of the source-side
// column from `base` in the bits above TC_KIND_SHIFT.  TC_KIND_VEC4 marks a 4-aligned column whose group of four is one
// aligned float4 on either side.
constexpr int TC_KIND_DPS = 6, TC_KIND_MASK = 127, TC_KIND_VEC4 = 128, TC_KIND_SHIFT = 8;

template <class Args>
__device__ __forceinline__ TcCol tc_match_col(const Args& a, int c) {
  TcCol t{nullptr, 0, -1};
  for (int si = 0; si < a.n_segs; ++si) {
    const Seg sg = a.segs[si];
    const int f = c - sg.row;
    if (f >= 0 && f < sg.width) {
      const float* b = a.arr[sg.arr] + sg.col + f;
      if (t.kind < 0) {
        t.base = b;
        t.ld = a.ld[sg.arr];
        t.kind = sg.kind;
      } else if (t.kind == SEG_DST && sg.kind == SEG_SRC && a.ld[sg.arr] == t.ld) {
        t.kind = TC_KIND_DPS | ((int)(b - t.base) << TC_KIND_SHIFT);
      }
    }
  }
  return t;
}

template <class Args>
__device__ __forceinline__ void tc_build_cols(const Args& a, TcCol* cols, int ncols, int tid, int nthreads) {
  for (int c = tid; c < ncols; c += nthreads) {
    TcCol t = tc_match_col(a, c);
    if ((c & 3) == 0 && c + 3 < ncols && (t.kind & TC_KIND_MASK) == TC_KIND_DPS && t.kind > 0) {
      const TcCol u = tc_match_col(a, c + 3);
      const int delta = t.kind >> TC_KIND_SHIFT;
      if (u.kind == t.kind && u.base == t.base + 3 && (t.ld & 3) == 0 && (delta & 3) == 0 &&
          (reinterpret_cast<uintptr_t>(t.base) & 15) == 0)
        t.kind |= TC_KIND_VEC4;
    }
    cols[c] = t;
  }
}

// A 16-column chunk of the node-phase input that is 16 consecutive, 16-byte aligned floats of ONE node array: loaded (and its
// cotangent stored) as 4 float4 per row instead of 16 table-driven scalar accesses -- with one row per lane every scalar
// access is its own 32-byte sector, so the load/store unit, not the memory, bounded those phases.
struct TcChunk {
  const float* base;  // array + first column of the chunk; nullptr: take the per-column path
  int ld;
};

template <class Args>
__device__ __forceinline__ void tc_build_chunks(const Args& a, TcChunk* chunks, int nchunks, int tid) {
  if (tid >= nchunks) return;
  TcChunk t{nullptr, 0};
  const int c = 16 * tid;
  int covering = 0;  // segments over these rows: more than one = a summed (hoisted) input, which takes the per-column path
  for (int si = 0; si < a.n_segs; ++si)
    if (a.segs[si].row < c + 16 && c < a.segs[si].row + a.segs[si].width) ++covering;
  for (int si = 0; si < a.n_segs && covering == 1; ++si) {
    const Seg sg = a.segs[si];
    const int f = c - sg.row;
    if (sg.kind != SEG_DST || f < 0 || f + 16 > sg.width) continue;
    const float* b = a.arr[sg.arr] + sg.col + f;
    if (((sg.col + f) & 3) == 0 && (a.ld[sg.arr] & 3) == 0 && (reinterpret_cast<uintptr_t>(a.arr[sg.arr]) & 15) == 0) {
      t.base = b;
      t.ld = a.ld[sg.arr];
    }
  }
  chunks[tid] = t;
}

// n (4 or 2) float4 of row `r` of an aligned chunk
template <int NV>
__device__ __forceinline__ void tc_load_chunk(const TcChunk ch, int r, int first, float* v) {
  const float4* p = reinterpret_cast<const float4*>(ch.base + (size_t)r * ch.ld + first);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 t = __ldg(p + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

template <bool SUMMED = true>  // SUMMED = false: a kernel variant that never sees summed (hoisted) columns leaves the branch out
__device__ __forceinline__ float tc_gather_col(const TcCol t, int s, int d, int p, int pg) {
  if (t.kind < 0) return 0.f;
  if (SUMMED && (t.kind & TC_KIND_MASK) == TC_KIND_DPS)
    return __ldg(t.base + (size_t)d * t.ld) + __ldg(t.base + (size_t)s * t.ld + (t.kind >> TC_KIND_SHIFT));
  const int i0 = (t.kind == SEG_SRC || t.kind == SEG_SMD) ? s : (t.kind == SEG_EDGE ? p : (t.kind == SEG_GRAPH ? pg : d));
  float v = t.base[(size_t)i0 * t.ld];
  if (t.kind == SEG_SMD) v -= t.base[(size_t)d * t.ld];
  if (t.kind == SEG_DMS) v -= t.base[(size_t)s * t.ld];
  return v;
}

// four columns of a TC_KIND_VEC4 group: one float4 from the destination's row plus one from the source's
__device__ __forceinline__ void tc_gather_dps4(const TcCol t, int s, int d, float* v) {
  const float4 x = __ldg(reinterpret_cast<const float4*>(t.base + (size_t)d * t.ld));
  const float4 y = __ldg(reinterpret_cast<const float4*>(t.base + (size_t)s * t.ld + (t.kind >> TC_KIND_SHIFT)));
  v[0] = x.x + y.x; v[1] = x.y + y.y; v[2] = x.z + y.z; v[3] = x.w + y.w;
}

// tanh(x) = 1 - 2 / (2^(2x log2 e) + 1) on the SFU exponential and reciprocal: 5 instructions, absolute error <= 1.5e-7
// over the whole range (saturates correctly to +-1), i.e. at the level of one float32 rounding of an O(1) activation.
__device__ __forceinline__ float tc_tanh(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
  return fmaf(-2.f, r, 1.f);
}

__device__ __noinline__ float tc_act_slow(int act, float x) { return act_fwd(act, x); }

// activation of a 16-value chunk with the dispatch hoisted out of the element loop (accurate float32 variants)
__device__ __forceinline__ void tc_act16(int act, float (&f)[16]) {
  switch (act) {
    case NGPDE_ACT_IDENTITY: break;
    case NGPDE_ACT_RELU:
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
      break;
    case NGPDE_ACT_TANH:
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = tc_tanh(f[j]);
      break;
    case NGPDE_ACT_SIGMOID:
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = __fdividef(1.f, 1.f + expf(-f[j]));
      break;
    case NGPDE_ACT_SWISH:
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = __fdividef(f[j], 1.f + expf(-f[j]));
      break;
    default:
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = tc_act_slow(act, f[j]);
      break;
  }
}

// the same for the 8-value chunks of the gather (hoisted first layer: the activation is applied to the gathered input)
__device__ __forceinline__ void tc_act8(int act, float (&f)[8]) {
  switch (act) {
    case NGPDE_ACT_IDENTITY: break;
    case NGPDE_ACT_RELU:
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      break;
    case NGPDE_ACT_TANH:
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = tc_tanh(f[j]);
      break;
    case NGPDE_ACT_SIGMOID:
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __fdividef(1.f, 1.f + expf(-f[j]));
      break;
    case NGPDE_ACT_SWISH:
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __fdividef(f[j], 1.f + expf(-f[j]));
      break;
    default:
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = tc_act_slow(act, f[j]);
      break;
  }
}

__device__ __forceinline__ void tc_split16(const float (&f)[16], uint32_t (&hi)[16], uint32_t (&lo)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float h = umma::tf32_hi(f[j]);
    hi[j] = __float_as_uint(h);
    lo[j] = __float_as_uint(umma::tf32_lo(f[j], h));
  }
}

// the "ones" block (1, 0, ..., 0) that meets the bias row of the next weight image
__device__ __forceinline__ void tc_store_ones(uint32_t tAhi, uint32_t tAlo, uint32_t col) {
  const uint32_t one[8] = {0x3f800000u, 0, 0, 0, 0, 0, 0, 0}, zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  umma::tmem_st8(tAhi + col, one);
  umma::tmem_st8(tAlo + col, zero);
}

// issue the 3xTF32 MMAs of one Dense layer: D[128 x Np] = A[128 x Kp] (TMEM hi/lo) x W (smem hi/lo images, MN-major)
__device__ __forceinline__ void tc_issue_layer(const TcLayout& lay, int l, uint32_t wblk_smem, uint32_t tD, uint32_t tAhi,
                                               uint32_t tAlo) {
  const uint32_t idesc = umma::make_idesc(TC_TILE, lay.Np[l], /*a_mn=*/0, /*b_mn=*/1);
  const uint32_t hi = wblk_smem + 4u * lay.img_off[l], lo = hi + 4u * lay.img_floats[l];
  const uint32_t lbo = 128u * lay.Kp[l];
  const uint64_t dhi = umma::make_sdesc(hi, lbo, 512, 1), dlo = umma::make_sdesc(lo, lbo, 512, 1);
  const int nks = lay.Kp[l] / 8;
  // one K-step = 8 rows of the image = 1024 bytes = 64 units of the descriptor's 16-byte address field.
  // Cross terms first, hi*hi last (see ngpde_umma.cuh).
  umma::mma_tf32_ts(tD, tAlo, dhi, idesc, 0);
#pragma unroll 4
  for (int ks = 1; ks < nks; ++ks) umma::mma_tf32_ts(tD, tAlo + ks * 8, dhi + (uint64_t)(ks * 64), idesc, 1);
#pragma unroll 4
  for (int ks = 0; ks < nks; ++ks) umma::mma_tf32_ts(tD, tAhi + ks * 8, dlo + (uint64_t)(ks * 64), idesc, 1);
#pragma unroll 4
  for (int ks = 0; ks < nks; ++ks) umma::mma_tf32_ts(tD, tAhi + ks * 8, dhi + (uint64_t)(ks * 64), idesc, 1);
}

template <bool NODE>
__global__ void __launch_bounds__(TC_THREADS, 1) mp_fwd_tc_kernel(const __grid_constant__ TcFwdArgs a) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bars[TC_GROUPS];
  __shared__ __align__(8) uint64_t wbar;
  __shared__ uint32_t tmem_slot;
  const TcLayout& lay = a.lay;
  const int tid = threadIdx.x, grp = umma::uniform_i32(threadIdx.x / TC_GTHREADS), gt = tid % TC_GTHREADS;
  const int gwarp = umma::uniform_i32((threadIdx.x % TC_GTHREADS) >> 5);  // warp index inside the group
  const int row = gt & 127;  // tile row this thread serves: lane quarter (gt>>5)&3, lane gt&31
  const int half = gt >> 7;  // which 16-column chunks of a layer's output it handles (chunk & 1 == half)
  float* wblk = reinterpret_cast<float*>(smem);
  TcCol* cols = reinterpret_cast<TcCol*>(smem + a.off_cols);
  TcChunk* chunks = reinterpret_cast<TcChunk*>(cols + lay.Kd[0]);
  float* M = reinterpret_cast<float*>(smem + a.off_groups + grp * a.group_bytes);  // [128][dout + 1]
  const int dout = a.dout, ldm = dout + 1, aggr = a.aggr;
  float* __restrict__ out = a.out;

  if (tid < 32) umma::tmem_alloc(&tmem_slot, lay.tmem_cols);
  if (tid == 0) {
    umma::mbar_init(&wbar, 1);
    for (int g = 0; g < TC_GROUPS; ++g) umma::mbar_init(&bars[g], 1);
    umma::fence_mbar_init();
  }
  tc_build_cols(a, cols, lay.Kd[0], tid, TC_THREADS);
  if (NODE) tc_build_chunks(a, chunks, lay.Kd[0] >> 4, tid);
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  if (tid == 0) {
    const uint32_t bytes = 4u * lay.block_floats;
    mbar_arrive_expect_tx(&wbar, bytes);
    for (uint32_t off = 0; off < bytes; off += 16384)
      bulk_g2s(smem + off, reinterpret_cast<const uint8_t*>(a.wblock) + off, min(16384u, bytes - off), &wbar);
  }
  umma::mbar_wait(&wbar, 0);

  const uint32_t tmem = umma::uniform_u32(tmem_slot);
  const uint32_t lane_addr = (uint32_t)(((gt >> 5) & 3) * 32) << 16;
  const uint32_t tD = tmem + grp * lay.cols_group, tAhi = tD + TC_MAXN, tAlo = tAhi + lay.kmax;
  const uint32_t wblk_smem = umma::smem_u32(wblk);
  uint32_t phase = 0;
  const int L = lay.L, Kd0 = lay.Kd[0], gdiv = a.tg.gdiv;
  const bool skip0 = a.skip_l0 && L > 1 && lay.Np[0] == Kd0;
  const float ident = aggr == NGPDE_AGGR_MAX ? -INFINITY : (aggr == NGPDE_AGGR_MIN ? INFINITY : 0.f);

  for (int unit = blockIdx.x * TC_GROUPS + grp; unit < a.tg.n_units; unit += gridDim.x * TC_GROUPS) {
    int n0, n1, kbeg, kend;
    if (NODE) {
      n0 = unit * TC_TILE;
      n1 = min(a.tg.N, n0 + TC_TILE);
      kbeg = n0;
      kend = n1;
    } else {
      n0 = a.tg.unit_ptr[unit];
      n1 = a.tg.unit_ptr[unit + 1];
      kbeg = a.tg.rowptr[n0];
      kend = a.tg.rowptr[n1];
      for (int item = gt; item < (n1 - n0) * dout; item += TC_GTHREADS) {  // isolated destinations
        const int jj = item / dout;
        if (a.tg.rowptr[n0 + jj] == a.tg.rowptr[n0 + jj + 1]) out[(size_t)n0 * dout + item] = ident;
      }
    }
    for (int k0 = kbeg; k0 < kend; k0 += TC_TILE) {
      const int ne = min(TC_TILE, kend - k0);
      const bool valid = row < ne;
      int s = 0, d = 0, p = 0;
      if (valid) {
        if (NODE) {
          s = d = p = k0 + row;
        } else {
          s = a.tg.src[k0 + row];
          d = a.tg.dst[k0 + row];
          p = a.tg.perm[k0 + row];
        }
      }
      const int pg = p / gdiv;
      // ---- gather this row's MLP input straight into TMEM (8-column chunks alternate between the two halves) ----
      for (int c0 = 8 * half; c0 < Kd0; c0 += 16) {
        uint32_t hi[8], lo[8];
        float vv[8];
        const TcChunk ch = NODE ? chunks[c0 >> 4] : TcChunk{nullptr, 0};
        if (NODE && ch.base != nullptr) {
#pragma unroll
          for (int j = 0; j < 8; ++j) vv[j] = 0.f;
          if (valid) tc_load_chunk<2>(ch, d, c0 & 15, vv);
        } else {
          const TcCol t0 = cols[c0], t4 = cols[c0 + 4];
          if ((t0.kind & TC_KIND_VEC4) && (t4.kind & TC_KIND_VEC4) && t0.kind > 0 && t4.kind > 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) vv[j] = 0.f;
            if (valid) {
              tc_gather_dps4(t0, s, d, vv);
              tc_gather_dps4(t4, s, d, vv + 4);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) vv[j] = valid ? tc_gather_col(cols[c0 + j], s, d, p, pg) : 0.f;
          }
        }
        if (skip0) tc_act8(a.act[0], vv);  // identity first layer: these ARE layer 1's input columns
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = vv[j];
          const float h = umma::tf32_hi(v);
          hi[j] = __float_as_uint(h);
          lo[j] = __float_as_uint(umma::tf32_lo(v, h));
        }
        umma::tmem_st8(tAhi + lane_addr + c0, hi);
        umma::tmem_st8(tAlo + lane_addr + c0, lo);
      }
      if (half == 0) tc_store_ones(tAhi + lane_addr, tAlo + lane_addr, Kd0);
      umma::tmem_wait_st();
      umma::tc_fence_before();
      group_bar(grp);

      for (int l = skip0 ? 1 : 0; l < L; ++l) {
        if (gwarp == 0) {
          // one elected lane issues the MMAs (elect.sync keeps the operands in uniform registers: no R2UR waterfall per
          // MMA) and waits for their completion; everybody else parks on barriers (a spinning try_wait loop in 255
          // threads would steal issue slots from the other group's epilogue)
          if (umma::elect_one_sync()) {
            umma::tc_fence_after();
            tc_issue_layer(lay, l, wblk_smem, tD, tAhi, tAlo);
            umma::mma_commit(&bars[grp]);
            umma::mbar_wait(&bars[grp], phase);
          }
          __syncwarp();
        }
        phase ^= 1;
        group_bar(grp);
        umma::tc_fence_after();
        const int act = a.act[l], Np = lay.Np[l];
        const bool last = l == L - 1;
        for (int c0 = 16 * half; c0 < Np; c0 += 32) {
          uint32_t v[16];
          umma::tmem_ld16(tD + lane_addr + c0, v);
          umma::tmem_wait_ld();
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]);
          tc_act16(act, f);
          if (!last) {
            uint32_t lo[16];
            tc_split16(f, v, lo);
            umma::tmem_st16(tAhi + lane_addr + c0, v);
            umma::tmem_st16(tAlo + lane_addr + c0, lo);
          } else if (NODE) {
            if (valid) {
              float* o = out + (size_t)(k0 + row) * a.out_ld + c0;
              if (c0 + 16 <= dout && (a.out_ld & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (c0 + j < dout) o[j] = f[j];
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < dout) M[row * ldm + c0 + j] = f[j];
          }
        }
        if (!last) {
          if (half == 0) tc_store_ones(tAhi + lane_addr, tAlo + lane_addr, Np);  // Np[l] == Kd[l+1]
          umma::tmem_wait_st();
          umma::tc_fence_before();
          group_bar(grp);
        }
      }
      if (!NODE) {
        // ---- ordered per-destination reduction of the tile's messages (ascending stored edge position) ----
        group_bar(grp);
        const int total = (n1 - n0) * dout;
        for (int item = gt; item < total; item += TC_GTHREADS) {
          const int jj = item / dout, c = item - jj * dout;
          const int j = n0 + jj;
          const int r0 = a.tg.rowptr[j], r1 = a.tg.rowptr[j + 1];
          const int lo = max(r0, k0), hi = min(r1, k0 + ne);
          if (lo >= hi) continue;
          float acc = (lo == r0) ? ident : out[(size_t)j * dout + c];
          const float* m = M + c + (lo - k0) * ldm;
          const int cnt = hi - lo;
          if (aggr == NGPDE_AGGR_MAX) {
            for (int e = 0; e < cnt; ++e) acc = fmaxf(acc, m[e * ldm]);
          } else if (aggr == NGPDE_AGGR_MIN) {
            for (int e = 0; e < cnt; ++e) acc = fminf(acc, m[e * ldm]);
          } else {
            for (int e = 0; e < cnt; ++e) acc = __fadd_rn(acc, m[e * ldm]);
            if (aggr == NGPDE_AGGR_MEAN && hi == r1) acc = __fdiv_rn(acc, (float)(r1 - r0));
          }
          out[(size_t)j * dout + c] = acc;
        }
        group_bar(grp);
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (tid < 32) umma::tmem_dealloc(tmem, lay.tmem_cols);
}

}  // namespace ngpde
