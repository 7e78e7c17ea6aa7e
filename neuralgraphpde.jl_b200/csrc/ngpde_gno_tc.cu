// tcgen05 / TMEM GEMM of the factored GNOConv evaluation (ngpde_gno.cuh): C[M][N] = op(A) op(B), FP32-accurate through
// the 3xTF32 split (ngpde_umma.cuh), FP32 accumulation in tensor memory.
//
// One CTA of 128 threads owns a 128 x 64 output tile.  Per k-block of 32 floats the A tile goes to TENSOR MEMORY (thread =
// row = TMEM lane; hi and lo halves of the TF32 split side by side -- an SS-form first version was bound by shared-memory
// bandwidth: 6 KB of operand reads per MMA plus the staging writes) and the B tile to shared memory as K-major
// SWIZZLE_128B images (rows = n, 128-byte rows of 32 k-values; hi and lo image each):
// global -> registers (prefetched two blocks ahead) -> TF32 split -> TMEM / shared memory; one elected thread then issues
// 4 k-steps x 3 products  lo*hi, hi*lo, hi*hi  of tcgen05.mma kind::tf32 (M = 128, N = 64, K = 8) and commits them to the
// stage's mbarrier; two stages, so the staging of block i+1 overlaps the MMAs of block i.  Sources stored with the
// OTHER index contiguous (B = [K][N] of mbar = S B; both operands of dB = S' DM) are transposed on the way in with a lane
// map (8 rows x 4 k per warp store) that is bank-conflict free under the swizzle and reads full 32-byte sectors.
// Epilogue: tcgen05.ld, thread = output row, 256 contiguous bytes per row.
//
// Operand conventions (K-major SWIZZLE_128B read with SBO = 1024, k-step = +32 bytes) are the ones pinned on hardware by
// tools/umma_probe.cu mode 1 img 0; tools/gemm_check.py checks every variant of this kernel against a float64 product.
#include "ngpde_gno.cuh"
#include "ngpde_umma.cuh"

namespace ngpde {
namespace {

using namespace umma;

constexpr int TBM = 128, TBN = 64, TBK = 32, TGT = 128;
constexpr int A_IMG = TBM * TBK;                     // floats per A image (16 KB)
constexpr int B_IMG = TBN * TBK;                     // floats per B image (8 KB)
constexpr int STAGE_FLOATS = 2 * B_IMG;              // B hi | B lo   (16 KB); the A operand lives in tensor memory
constexpr int TC_SMEM_BYTES = 2 * STAGE_FLOATS * 4 + TBM * TBK * 4 + 1024;  // + the warps' transpose patches
constexpr int TMEM_COLS = 256;                       // D0 | D1 (64 each) | stage 0: A hi, A lo (32 each) | stage 1: A hi, A lo
constexpr int COL_A = 2 * TBN;
constexpr int CHUNK = 8;                             // k-blocks (256 k-values) accumulated in TMEM before an FP32 flush

struct TcGemmArgs {
  const float* A;
  const float* B;
  float* C;
  const int* deg_rowptr;
  long long M, K;
  int N, lda, ldb, ldc;
  long long k_per_split;
};

// ---- global -> registers.  ROWS = 128 (A) or 64 (B).  NV values per thread = ROWS * 32 / 128.
// SRC_T == false: source [rows][K] (k contiguous): float4 along k; thread's v-th float4: row = tid/8 + 16 v, kq = tid % 8.
// SRC_T == true : source [K][rows] (row index contiguous): scalars; combo = warp + 4 v, row = 8 (combo / 8) + lane / 4,
//                 k = 4 (combo % 8) + lane % 4.
template <int ROWS, bool SRC_T>
__device__ __forceinline__ void tile_load(float (&r)[ROWS / 4], const float* __restrict__ P, int ld, long long row0,
                                          long long nrows, long long k0, long long kend, int tid) {
  const bool full = k0 + TBK <= kend && row0 + ROWS <= nrows;
  if (!SRC_T) {
    const float* p = P + (size_t)(row0 + (tid >> 3)) * ld + k0 + 4 * (tid & 7);
    if (full) {
#pragma unroll
      for (int v = 0; v < ROWS / 16; ++v) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(p + (size_t)(16 * v) * ld));
        r[4 * v + 0] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
      }
    } else {
#pragma unroll
      for (int v = 0; v < ROWS / 16; ++v) {
        const long long gr = row0 + (tid >> 3) + 16 * v, gk = k0 + 4 * (tid & 7);
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < nrows && gk < kend) x = __ldg(reinterpret_cast<const float4*>(p + (size_t)(16 * v) * ld));
        r[4 * v + 0] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
      }
    }
  } else {
    const int warp = tid >> 5, lane = tid & 31;
    // combo = warp + 4 v: row = 8 (combo / 8) + lane / 4 = 8 (v / 2) + lane / 4, k = 4 (combo % 8) + lane % 4
    const float* p = P + (size_t)(k0 + 4 * warp + (lane & 3)) * ld + row0 + (lane >> 2);
    const float* p16 = p + (size_t)16 * ld;
    if (full) {
#pragma unroll
      for (int v = 0; v < ROWS / 4; ++v) r[v] = __ldg(((v & 1) ? p16 : p) + 8 * (v >> 1));
    } else {
#pragma unroll
      for (int v = 0; v < ROWS / 4; ++v) {
        const int row = 8 * (v >> 1) + (lane >> 2), k = 4 * (warp + 4 * (v & 1)) + (lane & 3);
        r[v] = (row0 + row < nrows && k0 + k < kend) ? __ldg(((v & 1) ? p16 : p) + 8 * (v >> 1)) : 0.f;
      }
    }
  }
}

// ---- registers -> hi / lo images (K-major SWIZZLE_128B, rows of 32 floats)
template <int ROWS, bool SRC_T>
__device__ __forceinline__ void tile_store(const float (&r)[ROWS / 4], float* __restrict__ hi, float* __restrict__ lo,
                                           int tid) {
  if (!SRC_T) {
#pragma unroll
    for (int v = 0; v < ROWS / 16; ++v) {
      const int row = (tid >> 3) + 16 * v, kq = tid & 7;
      const uint32_t off = (uint32_t)(row * 32 + ((kq ^ (row & 7)) << 2));
      float4 h, l;
      h.x = tf32_hi(r[4 * v + 0]); l.x = tf32_lo(r[4 * v + 0], h.x);
      h.y = tf32_hi(r[4 * v + 1]); l.y = tf32_lo(r[4 * v + 1], h.y);
      h.z = tf32_hi(r[4 * v + 2]); l.z = tf32_lo(r[4 * v + 2], h.z);
      h.w = tf32_hi(r[4 * v + 3]); l.w = tf32_lo(r[4 * v + 3], h.w);
      *reinterpret_cast<float4*>(hi + off) = h;
      *reinterpret_cast<float4*>(lo + off) = l;
    }
  } else {
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int v = 0; v < ROWS / 4; ++v) {
      const int combo = warp + 4 * v;
      const int row = 8 * (combo >> 3) + (lane >> 2), k = 4 * (combo & 7) + (lane & 3);
      const uint32_t off = sw128_offset(0, ROWS, row, k);
      const float h = tf32_hi(r[v]);
      hi[off] = h;
      lo[off] = tf32_lo(r[v], h);
    }
  }
}

// ---- A operand: thread = tile row (= TMEM lane), 32 k-values per block.
// SRC_T == true : source [K][M]: 32 scalars of the thread's own row; a warp's lanes read 128 contiguous bytes per k.
// SRC_T == false: source [M][K]: read row-per-thread this would be 32 different 128-byte lines per warp instruction (ncu:
//                 L1 85 % busy, tensor pipe 17 %), so each warp reads ITS 32 rows coalesced (8 lanes x float4 per row,
//                 4 rows per instruction) and a_to_rows() turns the registers into row order through a warp-private,
//                 XOR-swizzled 4 KB shared-memory patch.
template <bool SRC_T>
__device__ __forceinline__ void a_load(float (&r)[TBK], const float* __restrict__ P, int ld, long long m0, long long M,
                                       long long k0, long long kend, int tid) {
  const bool full = k0 + TBK <= kend;
  if (!SRC_T) {
    const int warp = tid >> 5, lane = tid & 31;
    const long long kq = k0 + 4 * (lane & 7);
    const float* p = P + (size_t)(m0 + 32 * warp + (lane >> 3)) * ld + kq;
    const long long mrow = m0 + 32 * warp + (lane >> 3);
    if (full && m0 + TBM <= M) {
#pragma unroll
      for (int v = 0; v < TBK / 4; ++v) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(p + (size_t)(4 * v) * ld));
        r[4 * v + 0] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
      }
    } else {
#pragma unroll
      for (int v = 0; v < TBK / 4; ++v) {
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mrow + 4 * v < M && kq < kend) x = __ldg(reinterpret_cast<const float4*>(p + (size_t)(4 * v) * ld));
        r[4 * v + 0] = x.x; r[4 * v + 1] = x.y; r[4 * v + 2] = x.z; r[4 * v + 3] = x.w;
      }
    }
  } else {
    const long long m = m0 + tid;
    const float* p = P + (size_t)k0 * ld + m;
    if (m < M && full) {
#pragma unroll
      for (int k = 0; k < TBK; ++k) r[k] = __ldg(p + (size_t)k * ld);
    } else {
#pragma unroll
      for (int k = 0; k < TBK; ++k) r[k] = (m < M && k0 + k < kend) ? __ldg(p + (size_t)k * ld) : 0.f;
    }
  }
}
// registers in a_load's coalesced order -> the thread's own row (no-op for SRC_T); tw = this warp's 32 x 32 float patch
template <bool SRC_T>
__device__ __forceinline__ void a_to_rows(float (&r)[TBK], float* __restrict__ tw, int lane) {
  if (SRC_T) return;
#pragma unroll
  for (int v = 0; v < TBK / 4; ++v) {
    const int row = 4 * v + (lane >> 3), kq = lane & 7;
    *reinterpret_cast<float4*>(tw + row * 32 + ((kq ^ (row & 7)) << 2)) = make_float4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < TBK / 4; ++q) {
    const float4 x = *reinterpret_cast<const float4*>(tw + lane * 32 + ((q ^ (lane & 7)) << 2));
    r[4 * q + 0] = x.x; r[4 * q + 1] = x.y; r[4 * q + 2] = x.z; r[4 * q + 3] = x.w;
  }
  __syncwarp();
}
// TF32 split of the row into tensor memory: columns [col, col+32) = hi, [col+32, col+64) = lo
__device__ __forceinline__ void row_store_tmem(const float (&r)[TBK], uint32_t taddr) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float x = r[16 * h + q];
      const float xh = tf32_hi(x);
      hi[q] = __float_as_uint(xh);
      lo[q] = __float_as_uint(tf32_lo(x, xh));
    }
    tmem_st16(taddr + 16 * h, hi);
    tmem_st16(taddr + TBK + 16 * h, lo);
  }
}

template <bool A_T, bool B_T>
__global__ void __launch_bounds__(TGT) gno_gemm_tc_kernel(const TcGemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x;
  const int warp = uniform_i32(tid >> 5);
  const long long m0 = (long long)blockIdx.x * TBM;
  const int n0 = blockIdx.y * TBN;
  const long long kbeg = (long long)blockIdx.z * g.k_per_split;
  const long long kend = min(g.K, kbeg + g.k_per_split);
  const int nkb = kend > kbeg ? (int)((kend - kbeg + TBK - 1) / TBK) : 0;

  if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = uniform_u32(tmem_slot);
  const uint32_t smem_base = uniform_u32(smem_u32(smem));
  const uint32_t idesc = make_idesc(TBM, TBN, 0, 0);

  // FP32 accumulators of this thread's output row.  The tensor core truncates its FP32 accumulator at every MMA, so a long
  // K would drift (measured 3e-5 at K = 4160): every CHUNK k-blocks the partial sum leaves TMEM and is added here with a
  // rounded FP32 add, while the next chunk already accumulates into the other TMEM buffer (no pipeline bubble).
  float acc[TBN];
#pragma unroll
  for (int c = 0; c < TBN; ++c) acc[c] = 0.f;
  const uint32_t trow = tmem_d + ((uint32_t)((tid >> 5) * 32) << 16);
  auto flush = [&](int buf) {
#pragma unroll
    for (int c = 0; c < TBN; c += 16) {
      uint32_t v[16];
      tmem_ld16(trow + buf * TBN + c, v);
      tmem_wait_ld();
#pragma unroll
      for (int q = 0; q < 16; ++q) acc[c + q] += __uint_as_float(v[q]);
    }
    tc_fence_before();
  };

  // operand registers, prefetched two k-blocks ahead (the loop is unrolled by two so both sets are statically indexed)
  float ra0[TBK], rb0[TBN / 4], ra1[TBK], rb1[TBN / 4];
  auto fetch = [&](float (&ra)[TBK], float (&rb)[TBN / 4], int kb) {
    if (kb < nkb) {
      a_load<A_T>(ra, g.A, g.lda, m0, g.M, kbeg + (long long)kb * TBK, kend, tid);
      tile_load<TBN, B_T>(rb, g.B, g.ldb, n0, g.N, kbeg + (long long)kb * TBK, kend, tid);
    }
  };
  auto block = [&](float (&ra)[TBK], float (&rb)[TBN / 4], int kb) {
    const int s = kb & 1;
    const int buf = (kb / CHUNK) & 1;
    const bool chunk_start = (kb % CHUNK) == 0;
    float* st = smem + s * STAGE_FLOATS;
    if (kb >= 2) mbar_spin(&bar[s], (uint32_t)(((kb >> 1) - 1) & 1));  // the MMAs that read this stage have completed
    a_to_rows<A_T>(ra, smem + 2 * STAGE_FLOATS + (tid >> 5) * (32 * TBK), tid & 31);
    row_store_tmem(ra, trow + COL_A + s * 2 * TBK);
    tile_store<TBN, B_T>(rb, st, st + B_IMG, tid);
    fetch(ra, rb, kb + 2);
    tmem_wait_st();
    tc_fence_before();
    fence_async_smem();
    __syncthreads();
    if (warp == 0) {
      if (elect_one_sync()) {
        tc_fence_after();
        const uint32_t a_hi = tmem_d + (uint32_t)(COL_A + s * 2 * TBK);
        const uint32_t a_lo = a_hi + TBK;
        const uint32_t b_hi = smem_base + (uint32_t)(s * STAGE_FLOATS) * 4u;
        const uint32_t b_lo = b_hi + B_IMG * 4u;
        const uint32_t d = tmem_d + (uint32_t)(buf * TBN);
#pragma unroll
        for (int ks = 0; ks < TBK / 8; ++ks) {
          const uint64_t dbh = make_sdesc(b_hi + ks * 32, 0, 1024, 2), dbl = make_sdesc(b_lo + ks * 32, 0, 1024, 2);
          mma_tf32_ts(d, a_lo + ks * 8, dbh, idesc, (!chunk_start || ks > 0) ? 1 : 0);  // small cross terms first
          mma_tf32_ts(d, a_hi + ks * 8, dbl, idesc, 1);
          mma_tf32_ts(d, a_hi + ks * 8, dbh, idesc, 1);
        }
        mma_commit(&bar[s]);
      }
      __syncwarp();
    }
    if (chunk_start && kb > 0) {
      // the previous chunk ended with block kb-1: its commit covers every MMA of that chunk
      mbar_spin(&bar[(kb - 1) & 1], (uint32_t)(((kb - 1) >> 1) & 1));
      tc_fence_after();
      flush(buf ^ 1);
    }
  };
  fetch(ra0, rb0, 0);
  fetch(ra1, rb1, 1);
  for (int kb = 0; kb < nkb; kb += 2) {
    block(ra0, rb0, kb);
    if (kb + 1 < nkb) block(ra1, rb1, kb + 1);
  }

  // ---- epilogue: thread = output row ----
  float* C = g.C + (size_t)blockIdx.z * (size_t)g.M * g.ldc;
  const long long m = m0 + tid;
  float den = 1.f;
  bool zero = false;
  if (g.deg_rowptr != nullptr && m < g.M) {
    const int deg = g.deg_rowptr[m + 1] - g.deg_rowptr[m];
    den = (float)deg;
    zero = deg == 0;
  }
  if (nkb > 0) {
    const int last = nkb - 1;
    mbar_spin(&bar[last & 1], (uint32_t)((last >> 1) & 1));  // MMAs complete in issue order: the last commit covers all
    tc_fence_after();
    flush((last / CHUNK) & 1);
  }
  if (m < g.M) {
#pragma unroll
    for (int c = 0; c < TBN; c += 4) {
      const int n = n0 + c;
      if (n >= g.N) continue;  // N % 4 == 0
      float4 o = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
      if (zero) {
        o = make_float4(0.f, 0.f, 0.f, 0.f);
      } else if (g.deg_rowptr != nullptr) {
        o.x = __fdiv_rn(o.x, den); o.y = __fdiv_rn(o.y, den); o.z = __fdiv_rn(o.z, den); o.w = __fdiv_rn(o.w, den);
      }
      *reinterpret_cast<float4*>(C + (size_t)m * g.ldc + n) = o;
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

// ---- short-K, wide-N variant (T = DM B': K = out_chs <= 64, N = (K+1)*in_chs): per (128 x 64) tile the generic kernel
// would spend its time on set-up and on re-staging A, so here one CTA keeps its A tile (all of K) in tensor memory and walks
// the n-tiles: B tile j+1 is staged while the MMAs of tile j run, and tile j-1 leaves TMEM (two accumulators) for global
// memory underneath them.  A: [M][K], B: [N][K], both k-contiguous; KB = k-blocks of 32.
template <int KB>
__global__ void __launch_bounds__(TGT) gno_gemm_tc_nloop_kernel(const TcGemmArgs g) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar[2];
  constexpr int NSTAGE = KB * 2 * B_IMG;  // floats per stage: per k-block  B hi | B lo
  const int tid = threadIdx.x;
  const int warp = uniform_i32(tid >> 5);
  const long long m0 = (long long)blockIdx.x * TBM;
  const int ntiles = (g.N + TBN - 1) / TBN;

  if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = uniform_u32(tmem_slot);
  const uint32_t smem_base = uniform_u32(smem_u32(smem));
  const uint32_t idesc = make_idesc(TBM, TBN, 0, 0);
  const uint32_t trow = tmem_d + ((uint32_t)((tid >> 5) * 32) << 16);

  // A tile -> tensor memory, once
  {
    float ra[TBK];
#pragma unroll
    for (int kbi = 0; kbi < KB; ++kbi) {
      a_load<false>(ra, g.A, g.lda, m0, g.M, 32 * kbi, g.K, tid);
      a_to_rows<false>(ra, smem + 2 * NSTAGE + (tid >> 5) * (32 * TBK), tid & 31);
      row_store_tmem(ra, trow + COL_A + kbi * 2 * TBK);
    }
    tmem_wait_st();
    tc_fence_before();
  }

  float rb0[KB][TBN / 4], rb1[KB][TBN / 4];
  auto fetch = [&](float (&rb)[KB][TBN / 4], int j) {
    if (j < ntiles) {
#pragma unroll
      for (int kbi = 0; kbi < KB; ++kbi) tile_load<TBN, false>(rb[kbi], g.B, g.ldb, (long long)j * TBN, g.N, 32 * kbi, g.K, tid);
    }
  };
  // epilogue of tile t: TMEM -> registers (thread = row) -> warp-private swizzled patch -> coalesced stores (a row's 256
  // bytes by 16 lanes; written row-per-thread the stores are 32 half-filled sectors per instruction: 2.7 -> 1.9 ms at C4)
  float* patch = smem + 2 * NSTAGE + (tid >> 5) * (32 * TBN);  // shares its memory with the prologue's A patches
  const int lane = tid & 31;
  auto epilogue = [&](int t) {
    const int sb = t & 1;
    mbar_spin(&bar[sb], (uint32_t)((t >> 1) & 1));
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < TBN; c += 16) {
      uint32_t v[16];
      tmem_ld16(trow + sb * TBN + c, v);
      tmem_wait_ld();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int chunk = (c >> 2) + q;
        *reinterpret_cast<float4*>(patch + lane * TBN + ((chunk ^ (lane & 7)) << 2)) =
            make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
      }
    }
    tc_fence_before();
    __syncwarp();
#pragma unroll
    for (int v = 0; v < 16; ++v) {
      const int row = 2 * v + (lane >> 4), chunk = lane & 15;
      const float4 x = *reinterpret_cast<const float4*>(patch + row * TBN + ((chunk ^ (row & 7)) << 2));
      const long long mm = m0 + 32 * (tid >> 5) + row;
      const int n = t * TBN + 4 * chunk;
      if (mm < g.M && n < g.N) *reinterpret_cast<float4*>(g.C + (size_t)mm * g.ldc + n) = x;  // N % 4 == 0
    }
    __syncwarp();
  };
  auto tile = [&](float (&rb)[KB][TBN / 4], int j) {
    const int s = j & 1;
    float* st = smem + s * NSTAGE;
    if (j >= 2) mbar_spin(&bar[s], (uint32_t)(((j >> 1) - 1) & 1));  // tile j-2 done with this stage (its epilogue ran too)
#pragma unroll
    for (int kbi = 0; kbi < KB; ++kbi) tile_store<TBN, false>(rb[kbi], st + kbi * 2 * B_IMG, st + kbi * 2 * B_IMG + B_IMG, tid);
    fetch(rb, j + 2);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      if (elect_one_sync()) {
        tc_fence_after();
        const uint32_t d = tmem_d + (uint32_t)(s * TBN);
#pragma unroll
        for (int kbi = 0; kbi < KB; ++kbi) {
          const uint32_t a_hi = tmem_d + (uint32_t)(COL_A + kbi * 2 * TBK), a_lo = a_hi + TBK;
          const uint32_t b_hi = smem_base + (uint32_t)(s * NSTAGE + kbi * 2 * B_IMG) * 4u, b_lo = b_hi + B_IMG * 4u;
#pragma unroll
          for (int ks = 0; ks < TBK / 8; ++ks) {
            const uint64_t dbh = make_sdesc(b_hi + ks * 32, 0, 1024, 2), dbl = make_sdesc(b_lo + ks * 32, 0, 1024, 2);
            mma_tf32_ts(d, a_lo + ks * 8, dbh, idesc, (kbi > 0 || ks > 0) ? 1 : 0);
            mma_tf32_ts(d, a_hi + ks * 8, dbl, idesc, 1);
            mma_tf32_ts(d, a_hi + ks * 8, dbh, idesc, 1);
          }
        }
        mma_commit(&bar[s]);
      }
      __syncwarp();
    }
    if (j >= 1) epilogue(j - 1);
  };
  fetch(rb0, 0);
  fetch(rb1, 1);
  for (int j = 0; j < ntiles; j += 2) {
    tile(rb0, j);
    if (j + 1 < ntiles) tile(rb1, j + 1);
  }
  epilogue(ntiles - 1);
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

template <int KB>
int launch_nloop(const TcGemmArgs& g, cudaStream_t st) {
  constexpr int bytes = (2 * KB * 2 * B_IMG + TBM * TBN) * 4 + 1024;  // stages | A patch (prologue) = epilogue patch: 97 KB, two CTAs per SM
  static bool configured = false;
  if (!configured) {
    NGPDE_CUDA_TRY(cudaFuncSetAttribute(gno_gemm_tc_nloop_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    configured = true;
  }
  gno_gemm_tc_nloop_kernel<KB><<<(unsigned)((g.M + TBM - 1) / TBM), TGT, bytes, st>>>(g);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

template <bool A_T, bool B_T>
int launch(const TcGemmArgs& g, dim3 grid, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    NGPDE_CUDA_TRY(cudaFuncSetAttribute(gno_gemm_tc_kernel<A_T, B_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    configured = true;
  }
  gno_gemm_tc_kernel<A_T, B_T><<<grid, TGT, TC_SMEM_BYTES, st>>>(g);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

}  // namespace

bool gno_gemm_tc_supported(int lda, bool a_kmajor, int ldb, bool b_kmajor, int ldc, int N, int64_t K) {
  if ((ldc & 3) || (N & 3)) return false;
  if (!a_kmajor && ((lda & 3) || (K & 3))) return false;  // float4 reads along k
  if (!b_kmajor && ((ldb & 3) || (K & 3))) return false;
  return true;
}

int gno_gemm_tc(const float* A, int lda, bool a_kmajor, const float* B, int ldb, bool b_kmajor, float* C, int ldc, int64_t M,
                int N, int64_t K, int splits, const int* deg_rowptr, cudaStream_t st) {
  if (M <= 0 || N <= 0) return NGPDE_OK;
  NGPDE_REQUIRE(gno_gemm_tc_supported(lda, a_kmajor, ldb, b_kmajor, ldc, N, K), "gno_gemm_tc: unsupported strides");
  TcGemmArgs g;
  g.A = A; g.B = B; g.C = C; g.deg_rowptr = deg_rowptr;
  g.M = M; g.K = K; g.N = N; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  splits = splits < 1 ? 1 : splits;
  long long kps = (K + splits - 1) / splits;
  kps = (kps + TBK - 1) / TBK * TBK;
  g.k_per_split = kps > 0 ? kps : TBK;
  if (!a_kmajor && !b_kmajor && K <= 64 && splits == 1 && deg_rowptr == nullptr && N > TBN)
    return K <= 32 ? launch_nloop<1>(g, st) : launch_nloop<2>(g, st);
  dim3 grid((unsigned)((M + TBM - 1) / TBM), (unsigned)((N + TBN - 1) / TBN), (unsigned)splits);
  if (a_kmajor && b_kmajor) return launch<true, true>(g, grid, st);
  if (!a_kmajor && b_kmajor) return launch<false, true>(g, grid, st);
  if (!a_kmajor && !b_kmajor) return launch<false, false>(g, grid, st);
  return launch<true, false>(g, grid, st);
}

}  // namespace ngpde
