// Tensor-core (tcgen05 + TMEM) variant of the fused message-passing kernels for MLPs whose layers are at most 64 wide.
//
// One thread owns one edge (edge phase) or one node (node phase): a group of 128 threads owns a tile of 128 rows, which
// is the M dimension of every MMA.  Activations never leave the SM and never touch shared memory:
//     gather -> registers -> TMEM (A operand, lane = row, column = feature, split hi/lo for 3xTF32)
//     tcgen05.mma  D[128 x N](TMEM) = A(TMEM) x W(smem image)           3 passes: hi*hi + lo*hi + hi*lo, FP32 accumulate
//     tcgen05.ld D -> registers: + bias, activation (accurate tanhf/expf), split -> tcgen05.st -> next layer's A
// The last layer's rows go through a padded shared-memory tile only to be reduced per destination node in stored edge
// order (same atomic-free sequential reduction as the FFMA engine), or straight to global memory in the node phase.
// Weights are prepared once per call as SWIZZLE_128B_BASE32B images (tc_prep_weights_kernel) and pulled into shared
// memory with one TMA bulk copy per CTA; two 128-thread groups per CTA keep two tiles in flight so one group's
// epilogue overlaps the other's MMAs.  Operand conventions are pinned by tools/umma_probe.cu.
#pragma once
#include "ngpde_conv.cuh"
#include "ngpde_umma.cuh"

namespace ngpde {

constexpr int TC_TILE = 128;   // rows per tile = MMA M
constexpr int TC_GROUPS = 2;   // 128-thread groups per CTA
constexpr int TC_MAXN = 64;    // widest layer output the path accepts
constexpr int TC_THREADS = TC_TILE * TC_GROUPS;

struct TcLayout {
  int L;
  int K[NGPDE_MAX_LAYERS], N[NGPDE_MAX_LAYERS];    // logical layer shapes
  int Kp[NGPDE_MAX_LAYERS], Np[NGPDE_MAX_LAYERS];  // padded to multiples of 16
  int img_off[NGPDE_MAX_LAYERS];                   // float offset of the hi image; the lo image follows it
  int img_floats[NGPDE_MAX_LAYERS];                // floats of one image = ceil(Np/32) * Kp * 32
  int bias_off;                                    // float offset of biases [L][TC_MAXN]
  int block_floats;                                // images + biases
  int kmax;                                        // widest padded input
  int cols_group;                                  // TMEM columns per group: TC_MAXN (D) + 2*kmax (A hi, A lo)
  int tmem_cols;                                   // allocation: power of two >= 32
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
  float* out;           // edge phase: mbar [N][dout]; node phase: y [N][dout]
  int off_groups;       // byte offset of the per-group regions in dynamic shared memory
  int group_bytes;
};

// ---- prepared weight block: per layer hi image, lo image (SWIZZLE_128B_BASE32B, rows = k), then the biases ----
__global__ void tc_prep_weights_kernel(const float* __restrict__ params, MlpDev mlp, TcLayout lay, float* __restrict__ out) {
  const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int l = 0; l < lay.L; ++l) {
    const int K = lay.K[l], N = lay.N[l], Kp = lay.Kp[l];
    const int cols = lay.img_floats[l] / Kp;  // 32 * groups
    const float* W = params + mlp.w_off[l];
    float* hi = out + lay.img_off[l];
    float* lo = hi + lay.img_floats[l];
    for (int i = t0; i < Kp * cols; i += stride) {
      const int k = i / cols, n = i - k * cols;
      const float w = (k < K && n < N) ? W[(size_t)k * N + n] : 0.f;
      const float h = umma::tf32_hi(w);
      const uint32_t off = umma::sw128b32_offset(n >> 5, Kp, k, n & 31);
      hi[off] = h;
      lo[off] = umma::tf32_hi(w - h);
    }
    for (int n = t0; n < TC_MAXN; n += stride)
      out[lay.bias_off + l * TC_MAXN + n] = (mlp.b_off[l] >= 0 && n < N) ? params[mlp.b_off[l] + n] : 0.f;
  }
}

__device__ __forceinline__ void group_bar(int grp) { asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(TC_TILE) : "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   umma::smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(umma::smem_u32(bar))
               : "memory");
}

// value of input column `col` of the MLP for the item whose (src, dst, original position) are (s, d, p)
template <class Args>
__device__ __forceinline__ float tc_gather_value(const Args& a, int col, int s, int d, int p) {
  float v = 0.f;
  for (int si = 0; si < a.n_segs; ++si) {
    const Seg sg = a.segs[si];
    const int f = col - sg.row;
    if (f < 0 || f >= sg.width) continue;
    const float* __restrict__ A = a.arr[sg.arr] + sg.col + f;
    const size_t ld = (size_t)a.ld[sg.arr];
    switch (sg.kind) {
      case SEG_DST: v = A[d * ld]; break;
      case SEG_SRC: v = A[s * ld]; break;
      case SEG_SMD: v = A[s * ld] - A[d * ld]; break;
      case SEG_DMS: v = A[d * ld] - A[s * ld]; break;
      case SEG_EDGE: v = A[p * ld]; break;
      default: v = A[(size_t)(p / a.tg.gdiv) * ld]; break;  // SEG_GRAPH
    }
    break;
  }
  return v;
}

// issue the 3xTF32 MMAs of one Dense layer: D[128 x Np] = A[128 x Kp] (TMEM hi/lo) x W (smem hi/lo images, MN-major)
__device__ __forceinline__ void tc_issue_layer(const TcLayout& lay, int l, uint32_t wblk_smem, uint32_t tD, uint32_t tAhi,
                                               uint32_t tAlo) {
  const uint32_t idesc = umma::make_idesc(TC_TILE, lay.Np[l], /*a_mn=*/0, /*b_mn=*/1);
  const uint32_t hi = wblk_smem + 4u * lay.img_off[l], lo = hi + 4u * lay.img_floats[l];
  const uint32_t lbo = 128u * lay.Kp[l];
  const uint64_t dhi = umma::make_sdesc(hi, lbo, 512, 1), dlo = umma::make_sdesc(lo, lbo, 512, 1);
  const int nks = lay.Kp[l] / 8;
  // one K-step = 8 rows of the image = 1024 bytes = 64 units of the descriptor's 16-byte address field
  umma::mma_tf32_ts(tD, tAhi, dhi, idesc, 0);
#pragma unroll 4
  for (int ks = 1; ks < nks; ++ks) umma::mma_tf32_ts(tD, tAhi + ks * 8, dhi + (uint64_t)(ks * 64), idesc, 1);
#pragma unroll 4
  for (int ks = 0; ks < nks; ++ks) umma::mma_tf32_ts(tD, tAlo + ks * 8, dhi + (uint64_t)(ks * 64), idesc, 1);
#pragma unroll 4
  for (int ks = 0; ks < nks; ++ks) umma::mma_tf32_ts(tD, tAhi + ks * 8, dlo + (uint64_t)(ks * 64), idesc, 1);
}

template <bool NODE>
__global__ void __launch_bounds__(TC_THREADS, 1) mp_fwd_tc_kernel(const TcFwdArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t bars[TC_GROUPS];
  __shared__ __align__(8) uint64_t wbar;
  __shared__ uint32_t tmem_slot;
  const TcLayout& lay = a.lay;
  const int tid = threadIdx.x, grp = tid >> 7, gt = tid & 127;
  float* wblk = reinterpret_cast<float*>(smem);
  float* M = reinterpret_cast<float*>(smem + a.off_groups + grp * a.group_bytes);  // [128][dout + 1]
  const int ldm = a.dout + 1;

  if (tid < 32) umma::tmem_alloc(&tmem_slot, lay.tmem_cols);
  if (tid == 0) {
    umma::mbar_init(&wbar, 1);
    for (int g = 0; g < TC_GROUPS; ++g) umma::mbar_init(&bars[g], 1);
    umma::fence_mbar_init();
  }
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

  const uint32_t tmem = tmem_slot;
  const uint32_t lane_addr = (uint32_t)((gt >> 5) * 32) << 16;
  const uint32_t tD = tmem + grp * lay.cols_group, tAhi = tD + TC_MAXN, tAlo = tAhi + lay.kmax;
  const uint32_t wblk_smem = umma::smem_u32(wblk);
  const float* bias_all = wblk + lay.bias_off;
  uint32_t phase = 0;
  const int L = lay.L;

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
      const float ident = a.aggr == NGPDE_AGGR_MAX ? -INFINITY : (a.aggr == NGPDE_AGGR_MIN ? INFINITY : 0.f);
      for (int item = gt; item < (n1 - n0) * a.dout; item += TC_TILE) {
        const int jj = item / a.dout;
        if (a.tg.rowptr[n0 + jj] == a.tg.rowptr[n0 + jj + 1]) a.out[(size_t)n0 * a.dout + item] = ident;
      }
    }
    for (int k0 = kbeg; k0 < kend; k0 += TC_TILE) {
      const int ne = min(TC_TILE, kend - k0);
      const bool valid = gt < ne;
      int s = 0, d = 0, p = 0;
      if (valid) {
        if (NODE) {
          s = d = p = k0 + gt;
        } else {
          s = a.tg.src[k0 + gt];
          d = a.tg.dst[k0 + gt];
          p = a.tg.perm[k0 + gt];
        }
      }
      // ---- gather this row's MLP input straight into TMEM ----
      for (int c0 = 0; c0 < lay.Kp[0]; c0 += 8) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = (valid && c0 + j < lay.K[0]) ? tc_gather_value(a, c0 + j, s, d, p) : 0.f;
          const float h = umma::tf32_hi(v);
          hi[j] = __float_as_uint(h);
          lo[j] = __float_as_uint(umma::tf32_hi(v - h));
        }
        umma::tmem_st8(tAhi + lane_addr + c0, hi);
        umma::tmem_st8(tAlo + lane_addr + c0, lo);
      }
      umma::tmem_wait_st();
      umma::tc_fence_before();
      group_bar(grp);

      for (int l = 0; l < L; ++l) {
        if (gt < 32) {
          // one elected lane issues; its warp-mates park on __syncwarp instead of spinning in try_wait next to it
          if (gt == 0) {
            umma::tc_fence_after();
            tc_issue_layer(lay, l, wblk_smem, tD, tAhi, tAlo);
            umma::mma_commit(&bars[grp]);
          }
          __syncwarp();
        }
        umma::mbar_wait(&bars[grp], phase);
        phase ^= 1;
        umma::tc_fence_after();
        const float* bias = bias_all + l * TC_MAXN;
        const int act = a.act[l];
        const bool last = l == L - 1;
        for (int c0 = 0; c0 < lay.Np[l]; c0 += 16) {
          uint32_t v[16];
          umma::tmem_ld16(tD + lane_addr + c0, v);
          umma::tmem_wait_ld();
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + bias[c0 + j];
          if (act != NGPDE_ACT_IDENTITY) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = act_fwd(act, f[j]);
          }
          if (!last) {
            uint32_t lo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float h = umma::tf32_hi(f[j]);
              v[j] = __float_as_uint(h);
              lo[j] = __float_as_uint(umma::tf32_hi(f[j] - h));
            }
            umma::tmem_st16(tAhi + lane_addr + c0, v);
            umma::tmem_st16(tAlo + lane_addr + c0, lo);
          } else if (NODE) {
            if (valid) {
              float* o = a.out + (size_t)(k0 + gt) * a.dout + c0;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < a.dout) o[j] = f[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < a.dout) M[gt * ldm + c0 + j] = f[j];
          }
        }
        if (!last) {
          umma::tmem_wait_st();
          umma::tc_fence_before();
          group_bar(grp);
        }
      }
      if (!NODE) {
        // ---- ordered per-destination reduction of the tile's messages (ascending stored edge position) ----
        group_bar(grp);
        const int dm = a.dout;
        const int total = (n1 - n0) * dm;
        for (int item = gt; item < total; item += TC_TILE) {
          const int jj = item / dm, c = item - jj * dm;
          const int j = n0 + jj;
          const int r0 = a.tg.rowptr[j], r1 = a.tg.rowptr[j + 1];
          const int lo = max(r0, k0), hi = min(r1, k0 + ne);
          if (lo >= hi) continue;
          float acc = (lo == r0) ? (a.aggr == NGPDE_AGGR_MAX ? -INFINITY : (a.aggr == NGPDE_AGGR_MIN ? INFINITY : 0.f))
                                 : a.out[(size_t)j * dm + c];
          const float* m = M + c - (size_t)k0 * ldm;
          if (a.aggr == NGPDE_AGGR_MAX) {
            for (int e = lo; e < hi; ++e) acc = fmaxf(acc, m[(size_t)e * ldm]);
          } else if (a.aggr == NGPDE_AGGR_MIN) {
            for (int e = lo; e < hi; ++e) acc = fminf(acc, m[(size_t)e * ldm]);
          } else {
            for (int e = lo; e < hi; ++e) acc = __fadd_rn(acc, m[(size_t)e * ldm]);
            if (a.aggr == NGPDE_AGGR_MEAN && hi == r1) acc = __fdiv_rn(acc, (float)(r1 - r0));
          }
          a.out[(size_t)j * dm + c] = acc;
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
