// Tensor-core (tcgen05 + TMEM) backward of the fused message-passing kernels for MLPs whose layers are at most 64 wide
// (the twin of ngpde_tc.cuh).  One CTA of 512 threads owns one tile of 128 rows (edges or nodes); thread (row, q) serves
// the 16-column chunk q of its row, so a layer's 64 columns are handled by four warps at once.
//
// Per tile, with everything on chip:
//   1. recompute the forward activations Z_1 .. Z_{L-1} (3xTF32 MMAs as in the forward); they stay in TMEM as FP32;
//   2. G_L = cotangent of the tile's MLP output (gathered from dmbar[dst] / deg for the edge phase, dy[row] for nodes);
//   3. for l = L-1 .. 0:
//        input gradient   dZ_l = G_{l+1} W_l^T     M = 128 rows, A = G (TMEM, hi/lo), B = the FORWARD weight image read
//                                                  K-major (same shared-memory bytes serve both directions);
//        weight gradient  dW_l^T = G_{l+1}^T Z_l   M = 64 (output features), K = 64 rows per half tile, both operands
//                                                  staged in shared memory as MN-major SWIZZLE_128B_BASE32B images;
//        bias gradient    db_l = column sums of G_{l+1}: warp butterfly over the 32 rows a warp owns, fixed order;
//        G_l = dZ_l .* act'(Z_l)                   in registers, from TMEM.
//      Each tile's dW^T block is read out of TMEM and added to per-thread FP32 register accumulators (so the running sum
//      is rounded once per tile in FP32, not inside the tensor core), written as this CTA's partial at kernel end and
//      reduced over CTAs in fixed order by reduce_partials_kernel: deterministic, no atomics.
//   4. dZ_0 goes back to the arrays it was gathered from: node phase -> dx_direct / dmbar rows; edge phase -> per-edge
//      source-side rows (desrc, reduced later over the transpose) and the sequential per-destination sum (dxdst).
// 3xTF32 ordering: the two small cross terms are issued before hi*hi so that the accumulator is still small while they
// are added (the tensor core truncates its FP32 accumulator; pinned by tools/umma_probe.cu, profiles/r01b_umma_probe.log).
#pragma once
#include "ngpde_tc.cuh"

namespace ngpde {


struct TcBwdArgs {
  TileGraph tg;
  const float* arr[ARR_COUNT];
  int ld[ARR_COUNT];
  int n_segs;
  Seg segs[8];
  TcLayout lay;
  int act[NGPDE_MAX_LAYERS];
  int w_off[NGPDE_MAX_LAYERS], b_off[NGPDE_MAX_LAYERS];
  int n_params;
  const float* wblock;
  int aggr;
  int dout;                // width of the MLP output (= N[L-1])
  const float* gout_ptr;   // edge: dmbar / dy [N][dout]; node: dy [N][gout_ld]
  int gout_ld;
  const float* yact;       // node phase: cotangent *= act'(y) on load (y [N][dout]); NULL: off
  int yact_kind;
  int src_c0, src_w;       // x columns with source-side cotangents; desrc is [E][src_w]
  int dst_c0, dst_w;       // x columns with destination-side cotangents (the others of dxdst are not written)
  int direct_src;          // source-side cotangent row = dZ_0 row (hoisted input): stored from registers
  int skip_w0;             // layer 0's weight gradient is not wanted (identity of a hoisted first layer): no staging, no MMAs
  float* dparams_partial;  // [gridDim.x][n_params]
  float* dx_direct;        // node phase
  float* dmbar;            // node phase
  float* dxdst;            // edge phase [N][dx]
  float* desrc;            // edge phase [E][dx]
  int dx;
  int need_dz0, has_dst_side;
  // TMEM columns
  int c_zs[NGPDE_MAX_LAYERS];  // FP32 copy of Z_l (l >= 1)
  int c_a, a_width, c_d, c_dw, c_d0, c_dw0, tmem_cols;  // a_width: columns of one A image (hi or lo)
  int dw_alt;              // > 0: layers alternate between the accumulators c_dw and c_dw + dw_alt (deferred collection)
  // shared memory byte offsets
  int off_cols, off_stage, off_dz;
  int nzh, nzl;            // 32-column groups of the staged Z hi / lo images
  // weight images in shared memory: layer l's hi image starts at float offset woff[l] (its lo image follows).  When the
  // block does not fit next to a FULL-tile staging buffer, layers stream_a and stream_b share one slot (same woff) and
  // are swapped in by TMA bulk copies twice per tile (see the kernel); -1 = everything resident.
  int woff[NGPDE_MAX_LAYERS];
  int stream_a, stream_b;
  int opt;                 // experiment switches (NGPDE_TCB_OPT)
  long long* dbg;          // optional phase timestamps (clock64) of CTA 0, thread 64: [tile][64]; nullptr = off
};

// float offset inside a staged image of ROWS rows: element (r, c)
template <int ROWS>
__device__ __forceinline__ uint32_t stage_off(int r, int c) {
  return umma::sw128b32_offset(c >> 5, ROWS, r, c & 31);
}

// write 16 consecutive columns [c0, c0+16) of row r into a staged hi image and lo image (c0 % 16 == 0).
// A quarter warp holds 8 consecutive rows; rows r and r + 4 share the swizzled 32-byte chunk position, so the two 16-byte
// halves of a chunk are written in opposite order by the rows with bit 2 set: every 128-bit store instruction then
// covers 8 distinct 16-byte bank groups (conflict-free instead of 2-way conflicted).
template <int ROWS>
__device__ __forceinline__ void stage_chunk(float* img_hi, float* img_lo, int r, int c0, const float (&f)[16]) {
  const bool swp = (r & 4) != 0;
#pragma unroll
  for (int b = 0; b < 2; ++b) {  // two 32-byte blocks
    const uint32_t o = stage_off<ROWS>(r, c0 + 8 * b);
    const uint32_t o1 = o + (swp ? 4u : 0u), o2 = o + (swp ? 0u : 4u);
    float x[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      x[j] = swp ? f[8 * b + 4 + j] : f[8 * b + j];
      x[4 + j] = swp ? f[8 * b + j] : f[8 * b + 4 + j];
    }
    float4 h0, h1, l0, l1;
    h0.x = umma::tf32_hi(x[0]); h0.y = umma::tf32_hi(x[1]); h0.z = umma::tf32_hi(x[2]); h0.w = umma::tf32_hi(x[3]);
    h1.x = umma::tf32_hi(x[4]); h1.y = umma::tf32_hi(x[5]); h1.z = umma::tf32_hi(x[6]); h1.w = umma::tf32_hi(x[7]);
    l0.x = umma::tf32_lo(x[0], h0.x); l0.y = umma::tf32_lo(x[1], h0.y);
    l0.z = umma::tf32_lo(x[2], h0.z); l0.w = umma::tf32_lo(x[3], h0.w);
    l1.x = umma::tf32_lo(x[4], h1.x); l1.y = umma::tf32_lo(x[5], h1.y);
    l1.z = umma::tf32_lo(x[6], h1.z); l1.w = umma::tf32_lo(x[7], h1.w);
    *reinterpret_cast<float4*>(img_hi + o1) = h0;
    *reinterpret_cast<float4*>(img_lo + o1) = l0;
    *reinterpret_cast<float4*>(img_hi + o2) = h1;
    *reinterpret_cast<float4*>(img_lo + o2) = l1;
  }
}

// MMAs [i0, i1) of the 3 * (Kd/8) that make up one recomputed Dense layer (pass = i / nks: lo*hi, hi*lo, hi*hi), into the
// accumulator tDpart.  Two issuing warps take one half each, into separate accumulators (one thread sustains only one MMA
// per ~65 cycles, tools/tmem_bench.cu); the bias is added by the epilogue, so no "ones" column is involved.
__device__ __forceinline__ void tcb_issue_fwd_part(const TcLayout& lay, int l, uint32_t wimg_smem, uint32_t tDpart, uint32_t tAhi,
                                                   uint32_t tAlo, int i0, int i1) {
  const uint32_t idesc = umma::make_idesc(TC_TILE, lay.Np[l], /*a_mn=*/0, /*b_mn=*/1);
  const uint32_t hi = wimg_smem, lo = hi + 4u * lay.img_floats[l];
  const uint32_t lbo = 128u * lay.Kp[l];
  const uint64_t dhi = umma::make_sdesc(hi, lbo, 512, 1), dlo = umma::make_sdesc(lo, lbo, 512, 1);
  const int nks = lay.Kd[l] / 8;
  int pass = i0 / nks, ks = i0 - pass * nks;
#pragma unroll 1
  for (int i = i0; i < i1; ++i) {
    const uint32_t ta = (pass == 0 ? tAlo : tAhi) + ks * 8;
    const uint64_t db = (pass == 1 ? dlo : dhi) + (uint64_t)(ks * 64);
    umma::mma_tf32_ts(tDpart, ta, db, idesc, i > i0);
    if (++ks == nks) { ks = 0; ++pass; }
  }
}

// dZ_l = G W_l^T: A = G hi/lo in TMEM [128 x Np_l], B = forward weight image of layer l read K-major, N = Kd_l columns.
// K-step ks covers columns [8 ks, 8 ks + 8) of the image rows: group ks / 4 (gstride bytes apart), 32-byte block ks % 4.
__device__ __forceinline__ void tcb_issue_dgrad(const TcLayout& lay, int l, uint32_t wimg_smem, uint32_t tD, uint32_t tAhi,
                                                uint32_t tAlo) {
  const uint32_t idesc = umma::make_idesc(TC_TILE, lay.Kd[l], 0, 0);
  const uint32_t hi = wimg_smem;
  const uint64_t dhi = umma::make_sdesc(hi, 0, 512, 1);
  const uint64_t dlo = dhi + (uint64_t)(lay.img_floats[l] >> 2);  // image sizes are multiples of 1 KB: no field overflow
  const uint32_t g16 = 8u * lay.Kp[l];                             // group stride in 16-byte units
  const int nks = lay.Np[l] / 8;
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t ta = pass == 0 ? tAlo : tAhi;
    const uint64_t db = pass == 1 ? dlo : dhi;
#pragma unroll 4
    for (int ks = 0; ks < nks; ++ks)
      umma::mma_tf32_ts(tD, ta + ks * 8, db + (uint64_t)((ks >> 2) * g16 + (ks & 3) * 2u), idesc, (pass | ks) != 0);
  }
}

// dW_l^T += G^T Z over one staged image of ROWS rows (a half tile or the full tile): M = 64, K = ROWS, N = Kd_l;
// one K-step = 8 staged rows = 1024 bytes = 64 descriptor address units
template <int ROWS>
__device__ __forceinline__ void tcb_issue_wgrad(const TcLayout& lay, int l, uint32_t ghi, uint32_t glo, uint32_t zhi,
                                                uint32_t zlo, uint32_t tDw, int accumulate) {
  const uint32_t id_full = umma::make_idesc(64, lay.Kd[l], 1, 1);
  const uint32_t id_data = id_full;
  const uint32_t lbo = 128u * ROWS;
  const uint64_t dgh = umma::make_sdesc(ghi, lbo, 512, 1), dgl = umma::make_sdesc(glo, lbo, 512, 1);
  const uint64_t dzh = umma::make_sdesc(zhi, lbo, 512, 1), dzl = umma::make_sdesc(zlo, lbo, 512, 1);
  constexpr int nks = ROWS / 8;
#pragma unroll
  for (int ks = 0; ks < nks; ++ks)
    umma::mma_tf32_ss(tDw, dgl + (uint64_t)(ks * 64), dzh + (uint64_t)(ks * 64), id_full, accumulate | (ks > 0));
#pragma unroll
  for (int ks = 0; ks < nks; ++ks) umma::mma_tf32_ss(tDw, dgh + (uint64_t)(ks * 64), dzl + (uint64_t)(ks * 64), id_data, 1);
#pragma unroll
  for (int ks = 0; ks < nks; ++ks) umma::mma_tf32_ss(tDw, dgh + (uint64_t)(ks * 64), dzh + (uint64_t)(ks * 64), id_full, 1);
}

__device__ __forceinline__ float tcb_act_grad_y(int act, float y) {
  switch (act) {
    case NGPDE_ACT_IDENTITY: return 1.f;
    case NGPDE_ACT_TANH: return fmaf(-y, y, 1.f);
    case NGPDE_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    default: return act_grad_y(act, y);
  }
}

// where column c of dZ_0 goes in the node phase
struct TcDst {
  float* base;
  int ld;
};

#ifdef NGPDE_TCB_STAMPS
#define TCB_DIAG 1
#else
#define TCB_DIAG 0
#endif

// One lane per warp polls the mbarrier, the others park on __syncwarp: 512 threads spinning on mbarrier.try_wait slow the
// MMA-issuing lanes down by 2x and more (measured with the phase stamps).
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int opt) {
  if ((threadIdx.x & 31) == 0) {
    if (TCB_DIAG && (opt & 2)) umma::mbar_wait(bar, parity);
    else umma::mbar_spin(bar, parity);  // test_wait poll: try_wait may suspend the lane and wake it ~200 cycles late
  }
  __syncwarp();
}

// Phase stamps and the polling diagnostics are a DEVELOPER build (-DNGPDE_TCB_STAMPS, tools/tcb_phases*.py): the two dozen
// predicated stamp sites cost the C3 edge backward 2 % (0.703 -> 0.689 ms), so the shipped kernel does not contain them.
#ifndef NGPDE_TCB_STAMPS
#define TCB_STAMP(slot) do { } while (0)
#define TCB_STAMP_ANY(slot) do { } while (0)
#else
#define TCB_STAMP(slot)                                                                          \
  do {                                                                                           \
    if (a.dbg != nullptr && blockIdx.x == 0 && tid == TCB_STAMP_TID && dbg_tile < 8) a.dbg[dbg_tile * 64 + (slot)] = clock64(); \
  } while (0)
#define TCB_STAMP_ANY(slot)                                                                      \
  do {                                                                                           \
    if (a.dbg != nullptr && blockIdx.x == 0 && dbg_tile < 8) a.dbg[dbg_tile * 64 + (slot)] = clock64(); \
  } while (0)
#endif
constexpr int TCB_STAMP_TID = 160;  // warp 5 (rows 32..63, chunk 1): a worker that never issues MMAs

// column sums of a 16-value chunk over the 32 rows of a warp (fixed butterfly order): afterwards every lane holds the
// total of column  8*bit4 + 4*bit3 + 2*bit2 + bit1  of its lane id (lanes 2i and 2i+1 hold the same column).
// Split into three stages so that the caller can put independent work between them: while MMAs stream their operands
// from shared memory a shuffle takes ~250 cycles instead of ~25 (same data path; measured with tools/mma_rate.cu), and the
// five dependent levels of the butterfly would otherwise cost > 1000 cycles of pure latency per layer.
__device__ __forceinline__ void tcb_colsum_a(const float (&g)[16], int lane, float (&w)[8]) {
  const bool b4 = lane & 16;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float send = b4 ? g[j] : g[8 + j], keep = b4 ? g[8 + j] : g[j];
    w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
}
__device__ __forceinline__ void tcb_colsum_b(const float (&w)[8], int lane, float (&y)[2]) {
  float x[4];
  const bool b3 = lane & 8, b2 = lane & 4;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float send = b3 ? w[j] : w[4 + j], keep = b3 ? w[4 + j] : w[j];
    x[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float send = b2 ? x[j] : x[2 + j], keep = b2 ? x[2 + j] : x[j];
    y[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
}
__device__ __forceinline__ float tcb_colsum_c(const float (&y)[2], int lane) {
  const bool b1 = lane & 2;
  const float send = b1 ? y[0] : y[1], keep = b1 ? y[1] : y[0];
  float z = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  z += __shfl_xor_sync(0xffffffffu, z, 1);
  return z;
}

// Warp roles: 16 WORKER warps (thread = (row, 16-column chunk) of the tile) and one dedicated ISSUER warp.  tcgen05.mma issue
// is back-pressured at the rate the tensor pipe executes (measured, tools/mma_order.cu: the issuing lane is held for the
// whole batch), so an issuing worker would stall its own share of the epilogues and, through the next barrier, everybody
// else's.  The issuer only waits for "operands ready" mbarriers (bar_fg: A operand in TMEM; bar_fs: staged images in shared
// memory; 16 arrivals = one per worker warp) and issues; the workers never wait for the issuer, only for MMA completion
// (bar_d / bar_w, armed by tcgen05.commit).
// FULL: the weight-gradient operands of a layer are staged for the whole 128-row tile at once (all 16 warps, all four
// schedulers busy, one MMA batch of K = 128 per layer) instead of in two 64-row halves.  The 128 KB staging buffer only
// fits next to the weight images when two of them share a slot (stream_a / stream_b).
// EXT: the variant with everything the hoisted first layers (and GCNConv's fused act') added -- identity layer 0 without MMAs,
// summed gather, wide cotangent scatter, act'(y) on the cotangent load.  The plain variant does not even contain that code: the
// kernel is instruction-cache sensitive, and carrying it cost the C3 edge backward 8 % (0.698 -> 0.758 ms).
template <bool NODE, bool FULL, bool EXT>
__global__ void __launch_bounds__(TCB_THREADS, 1) mp_bwd_tc_kernel(const __grid_constant__ TcBwdArgs a) {
  constexpr int NW = TCB_WORKERS;  // worker threads
  constexpr int ROWS = FULL ? TC_TILE : TCB_HALF;  // rows of one staged image
  constexpr int NH = TC_TILE / ROWS;               // staging passes per layer
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar_d, bar_w, wbar, bar_ring, bar_fg, bar_fs;
  __shared__ uint32_t tmem_slot;
  const TcLayout& lay = a.lay;
  const int tid = threadIdx.x, warp = umma::uniform_i32(threadIdx.x >> 5), lane = tid & 31;
  const int lq = warp & 3;   // TMEM lane quarter this warp may touch
  const int q = warp >> 2;   // 16-column chunk served by this warp
  const int row = lq * 32 + lane;
  const int c0 = 16 * q;
  float* wblk = reinterpret_cast<float*>(smem);
  TcCol* cols = reinterpret_cast<TcCol*>(smem + a.off_cols);
  TcDst* dcols = reinterpret_cast<TcDst*>(cols + lay.Kd[0]);
  TcChunk* chunks = reinterpret_cast<TcChunk*>(dcols + lay.Kd[0]);   // node phase: aligned 16-column chunks of the input ...
  TcChunk* dchunks = chunks + TC_MAX_CHUNKS;                          // ... and of where its cotangent goes
  float* st_base = reinterpret_cast<float*>(smem + a.off_stage);  // Z hi | Z lo | G hi | G lo, each nz (2 for G) groups
  float* st_ghi = st_base + 2 * a.nzh * ROWS * 32;
  float* st_glo = st_ghi + 2 * ROWS * 32;
  float* DZ = reinterpret_cast<float*>(smem + a.off_dz);  // edge phase: [128][Kd0 + 1]; later reused for the db exchange
  const int L = lay.L, Kd0 = lay.Kd[0];
  // edge phase: the gathered layer-0 input is parked in the (not yet used) dZ_0 tile instead of being gathered twice
  const bool skip_w0 = EXT && a.skip_w0;
  const bool direct_src = EXT && a.direct_src;
  const bool keep_z0 = !NODE && a.need_dz0 && L > 1 && !skip_w0;
  // layer 0 is the identity of a hoisted first layer (skip_w0) and a thread's 16-column chunk of G_0 is its chunk of dZ_0:
  // dZ_0 = G_0 needs no input-gradient MMAs at all
  const bool id0 = skip_w0 && L > 1 && Kd0 <= 64 && lay.Np[0] == Kd0;
  // ... and no recompute MMAs either: Z_1 = act_0(gathered input), when the columns a thread gathers are its own chunk
  const bool id0r = id0 && (NODE || (Kd0 >> 2) == 16);

  if (tid < 32) umma::tmem_alloc(&tmem_slot, a.tmem_cols);
  if (tid == 0) {
    umma::mbar_init(&wbar, 1);
    umma::mbar_init(&bar_d, 1);
    umma::mbar_init(&bar_w, 1);
    umma::mbar_init(&bar_ring, 1);
    umma::mbar_init(&bar_fg, NW / 32);
    umma::mbar_init(&bar_fs, NW / 32);
    umma::fence_mbar_init();
  }
  tc_build_cols(a, cols, Kd0, tid, TCB_THREADS);
  if (NODE) {
    for (int c = tid; c < Kd0; c += TCB_THREADS) {
      TcDst t{nullptr, 0};
      for (int si = 0; si < a.n_segs; ++si) {
        const Seg sg = a.segs[si];
        const int f = c - sg.row;
        if (f < 0 || f >= sg.width || sg.kind != SEG_DST) continue;
        if (sg.arr == ARR_X && a.dx_direct != nullptr) t = TcDst{a.dx_direct + sg.col + f, a.ld[ARR_X]};
        if (sg.arr == ARR_M && a.dmbar != nullptr) t = TcDst{a.dmbar + sg.col + f, a.ld[ARR_M]};
      }
      dcols[c] = t;
    }
    tc_build_chunks(a, chunks, Kd0 >> 4, tid);
    if (tid < (Kd0 >> 4)) {
      // the cotangent of an aligned chunk of x / mbar goes to the same columns of dx_direct / dmbar
      TcChunk t{nullptr, 0};
      const int c = 16 * tid;
      for (int si = 0; si < a.n_segs; ++si) {
        const Seg sg = a.segs[si];
        const int f = c - sg.row;
        if (sg.kind != SEG_DST || f < 0 || f + 16 > sg.width) continue;
        float* dstb = sg.arr == ARR_X ? a.dx_direct : (sg.arr == ARR_M ? a.dmbar : nullptr);
        if (dstb != nullptr && ((sg.col + f) & 3) == 0 && (a.ld[sg.arr] & 3) == 0 && (reinterpret_cast<uintptr_t>(dstb) & 15) == 0) {
          t.base = dstb + sg.col + f;
          t.ld = a.ld[sg.arr];
        }
      }
      dchunks[tid] = t;
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  // one layer's (hi, lo) image pair: global prepared block -> its shared-memory slot, completion on `bar`
  auto load_layer = [&](int l, uint64_t* bar) {
    const uint32_t bytes = 8u * lay.img_floats[l];
    mbar_arrive_expect_tx(bar, bytes);
    const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wblock + lay.img_off[l]);
    uint8_t* dst = reinterpret_cast<uint8_t*>(wblk + a.woff[l]);
    for (uint32_t off = 0; off < bytes; off += 16384) bulk_g2s(dst + off, src + off, min(16384u, bytes - off), bar);
  };
  if (tid == 0) {
    uint32_t bytes = 0;
    for (int l = 0; l < L; ++l)
      if (l != a.stream_b) bytes += 8u * lay.img_floats[l];
    mbar_arrive_expect_tx(&wbar, bytes);
    for (int l = 0; l < L; ++l) {
      if (l == a.stream_b) continue;  // shares stream_a's slot: swapped in during the tile
      const uint32_t lb = 8u * lay.img_floats[l];
      const uint8_t* src = reinterpret_cast<const uint8_t*>(a.wblock + lay.img_off[l]);
      uint8_t* dst = reinterpret_cast<uint8_t*>(wblk + a.woff[l]);
      for (uint32_t off = 0; off < lb; off += 16384) bulk_g2s(dst + off, src + off, min(16384u, lb - off), &wbar);
    }
  }
  umma::mbar_wait(&wbar, 0);

  const uint32_t tmem = umma::uniform_u32(tmem_slot);
  const uint32_t lane_addr = (uint32_t)(lq * 32) << 16;
  const uint32_t tAhi = tmem + a.c_a, tAlo = tAhi + a.a_width, tD = tmem + a.c_d, tDw = tmem + a.c_dw;
  const uint32_t wblk_smem = umma::smem_u32(wblk);
  const uint32_t s_ghi = umma::smem_u32(st_ghi), s_glo = umma::smem_u32(st_glo);
  uint32_t ph_d = 0, ph_w = 0;
  const int gdiv = a.tg.gdiv, aggr = a.aggr, dout = a.dout;
  const bool streaming = a.stream_a >= 0;
  int dbg_tile = 0;

  auto worker_sync = [] { asm volatile("bar.sync 1, %0;" ::"n"(TCB_WORKERS) : "memory"); };
  // this warp's part of a hand-off to the issuer: everything each lane wrote (TMEM stores waited for, shared-memory stores
  // fenced towards the async proxy by the caller) is ordered before lane 0's arrival
  auto arrive = [&](uint64_t* bar) {
    umma::tc_fence_before();
    __syncwarp();
    if (lane == 0) umma::mbar_arrive(bar);
  };

  if (warp == NW / 32) {
    // ================= ISSUER: one elected lane walks the same tile / layer sequence and issues every MMA batch =================
    if (umma::elect_one_sync()) {
      uint32_t pf = 0, ps = 0, pr = 0, pd = 0;  // pd: parity of bar_d's next completion (diagnosis only)
      for (int unit = blockIdx.x; unit < a.tg.n_units; unit += gridDim.x) {
        int kbeg, kend;
        if (NODE) {
          kbeg = unit * TC_TILE;
          kend = min(a.tg.N, kbeg + TC_TILE);
        } else {
          kbeg = a.tg.rowptr[a.tg.unit_ptr[unit]];
          kend = a.tg.rowptr[a.tg.unit_ptr[unit + 1]];
        }
        for (int k0 = kbeg; k0 < kend; k0 += TC_TILE) {
          for (int l = id0r ? 1 : 0; l < L - 1; ++l) {  // recompute
            const int nmma = 3 * (lay.Kd[l] / 8);
            umma::mbar_wait(&bar_fg, pf);
            pf ^= 1;
            umma::tc_fence_after();
            tcb_issue_fwd_part(lay, l, wblk_smem + 4u * a.woff[l], tD, tAhi, tAlo, 0, nmma);
            umma::mma_commit(&bar_d);
            pd ^= 1;
          }
          for (int l = L - 1; l >= 0; --l) {
            const bool do_dgrad = (l > 0 || a.need_dz0) && !(id0 && l == 0);
            const uint32_t tDl = (l == 0) ? tmem + a.c_d0 : tD;
            const uint32_t tDwl = (l == 0) ? tmem + a.c_dw0 : tDw + ((l & 1) ? a.dw_alt : 0);
            const bool swapped = streaming && (l == a.stream_a || l == a.stream_b);
            if (do_dgrad) {
              umma::mbar_wait(&bar_fg, pf);
              pf ^= 1;
              umma::tc_fence_after();
              if (swapped) umma::mbar_wait(&bar_ring, pr);
              TCB_STAMP_ANY(32 + 4 * l);
              tcb_issue_dgrad(lay, l, wblk_smem + 4u * a.woff[l], tDl, tAhi, tAlo);
              umma::mma_commit(&bar_d);
              TCB_STAMP_ANY(33 + 4 * l);
              if (TCB_DIAG && (a.opt & 4)) {  // diagnosis: when does this batch really complete?  (bar_d parity as the workers track it)
                umma::mbar_spin(&bar_d, pd);
                TCB_STAMP_ANY(48 + l);
              }
              pd ^= 1;
            }
            if (swapped) pr ^= 1;
            const int gz = (lay.Kd[l] + 31) >> 5;
            const uint32_t s_zhi = umma::smem_u32(st_base), s_zlo = s_zhi + 4u * gz * ROWS * 32;
            for (int h = 0; h < ((id0 && l == 0) ? 0 : NH); ++h) {  // identity layer 0: no weight-gradient batch, no hand-off
              umma::mbar_wait(&bar_fs, ps);
              ps ^= 1;
              umma::tc_fence_after();
              TCB_STAMP_ANY(34 + 4 * l);
              if (!(skip_w0 && l == 0)) tcb_issue_wgrad<ROWS>(lay, l, s_ghi, s_glo, s_zhi, s_zlo, tDwl, h);
              umma::mma_commit(&bar_w);  // (an empty batch completes with whatever was issued before it)
              TCB_STAMP_ANY(35 + 4 * l);
            }
          }
          ++dbg_tile;
        }
      }
    }
    __syncwarp();
  } else {
  // ================= WORKERS =================

  // register accumulators of dW^T: per layer 8 values of the 16-column chunk (k = c0 + 8*(lane>=16) + j, n = 16*lq +
  // lane%16) plus up to 2 values of the columns beyond 64 (k = 64 + 8*t + 2*q + (lane>=16)), and one bias-gradient value
  constexpr int NACC = FULL ? 8 : 10;  // FULL kernels only take layers with Kd <= 64: no columns beyond 64
  float dwacc[TCB_MAXL][NACC];
  float dbacc[TCB_MAXL];
#pragma unroll
  for (int l = 0; l < TCB_MAXL; ++l) {
    dbacc[l] = 0.f;
#pragma unroll
    for (int j = 0; j < NACC; ++j) dwacc[l][j] = 0.f;
  }
  const bool upper = lane >= 16;
  const int n_items = NODE ? a.tg.N : a.tg.rowptr[a.tg.N];  // edges in the edge phase

  for (int unit = blockIdx.x; unit < a.tg.n_units; unit += gridDim.x) {
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
      if (a.has_dst_side) {
        for (int item = tid; item < (n1 - n0) * a.dx; item += NW) {
          const int jj = item / a.dx;
          if (a.tg.rowptr[n0 + jj] == a.tg.rowptr[n0 + jj + 1]) a.dxdst[(size_t)n0 * a.dx + item] = 0.f;
        }
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
      if (!NODE) {  // the next tile of this CTA: pull its index lines towards L1 while this one is processed
        const int kn = (k0 + TC_TILE < kend) ? k0 + TC_TILE + row
                                             : (unit + (int)gridDim.x < a.tg.n_units ? a.tg.rowptr[a.tg.unit_ptr[unit + gridDim.x]] + row : -1);
        if (kn >= 0 && kn < n_items && (lane & 7) == 0 && q < 3) {
          const int* ip = q == 0 ? a.tg.src : (q == 1 ? a.tg.dst : a.tg.perm);
          asm volatile("prefetch.global.L1 [%0];" ::"l"(ip + kn));
        }
      }
      const int pg = p / gdiv;
      TCB_STAMP(0);
      // the cotangent rows are first needed after the recompute: pull their lines towards L1 now
      if (valid && c0 < lay.Np[L - 1])
        asm volatile("prefetch.global.L1 [%0];" ::"l"(a.gout_ptr + (size_t)(NODE ? (k0 + row) : d) * a.gout_ld + c0));
      // mean: the cotangent of a message is dmbar / deg -- formed as dmbar * rn(1 / deg) (one rounding more than the true
      // division, far inside the gradient tolerance; sixteen IEEE divisions per thread cost ~1,000 cycles per tile)
      float rdeg = 1.f;
      if (!NODE && aggr == NGPDE_AGGR_MEAN && valid) rdeg = __frcp_rn((float)(a.tg.rowptr[d + 1] - a.tg.rowptr[d]));

      // ---- 1. forward recompute: Z_1 .. Z_{L-1} (two issuing warps, accumulators tD and tDw, bias added here) ----
      if (L > 1 && id0r) {
        // identity first layer (hoisted): Z_1 = act_0(gathered input) for this thread's chunk -- no MMAs, no A operand for layer 0
        if (c0 < Kd0) {
          float z[16];
          const TcChunk ch = NODE ? chunks[c0 >> 4] : TcChunk{nullptr, 0};
          if (NODE && ch.base != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) z[j] = 0.f;
            if (valid) tc_load_chunk<4>(ch, d, 0, z);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const TcCol t0 = cols[c0 + 4 * i];
              float v4[4];
              if (EXT && (t0.kind & TC_KIND_VEC4) && t0.kind > 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v4[j] = 0.f;
                if (valid) tc_gather_dps4(t0, s, d, v4);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v4[j] = valid ? tc_gather_col<EXT>(cols[c0 + 4 * i + j], s, d, p, pg) : 0.f;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) z[4 * i + j] = v4[j];
            }
          }
          tc_act16(a.act[0], z);
          uint32_t v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(z[j]);
          umma::tmem_st16(tmem + a.c_zs[1] + lane_addr + c0, v);
          if (L > 2) {
            uint32_t lo[16];
            tc_split16(z, v, lo);
            umma::tmem_st16(tAhi + lane_addr + c0, v);
            umma::tmem_st16(tAlo + lane_addr + c0, lo);
          }
        }
        umma::tmem_wait_st();
        if (L > 2) arrive(&bar_fg);
      } else if (L > 1) {
        if (NODE) {
          // node phase: whole 16-column chunks (vector loads where the chunk is an aligned run of one array)
          for (int cc = c0; cc < Kd0; cc += 64) {
            float v[16];
            const TcChunk ch = chunks[cc >> 4];
            if (ch.base != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = 0.f;
              if (valid) tc_load_chunk<4>(ch, d, 0, v);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = valid ? tc_gather_col<EXT>(cols[cc + j], s, d, p, pg) : 0.f;
            }
            uint32_t hi[16], lo[16];
            tc_split16(v, hi, lo);
            umma::tmem_st16(tAhi + lane_addr + cc, hi);
            umma::tmem_st16(tAlo + lane_addr + cc, lo);
          }
        }
        // edge phase: the four chunk warps of a row share the gather: a quarter of the Kd0 columns each, four at a time
        const int gq = NODE ? 0 : Kd0 >> 2;
        for (int cc = q * gq; cc < (q + 1) * gq; cc += 4) {
          uint32_t hi[4], lo[4];
          float v4[4];
          const TcCol t0 = cols[cc];
          if (EXT && (t0.kind & TC_KIND_VEC4) && t0.kind > 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) v4[j] = 0.f;
            if (valid) tc_gather_dps4(t0, s, d, v4);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v4[j] = valid ? tc_gather_col<EXT>(cols[cc + j], s, d, p, pg) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float v = v4[j];
            if (keep_z0) DZ[row * (Kd0 + 1) + cc + j] = v;  // re-read by layer 0's weight-gradient staging
            const float h = umma::tf32_hi(v);
            hi[j] = __float_as_uint(h);
            lo[j] = __float_as_uint(umma::tf32_lo(v, h));
          }
          umma::tmem_st4(tAhi + lane_addr + cc, hi);
          umma::tmem_st4(tAlo + lane_addr + cc, lo);
        }
        umma::tmem_wait_st();
        arrive(&bar_fg);
      }
      if (L > 1) {
#pragma unroll 1
        for (int l = id0r ? 1 : 0; l < L - 1; ++l) {
          const int Np = lay.Np[l];
          // bias row of the weight image (row Kd, unswizzled because Kd % 4 == 0): hi + lo.  Loaded BEFORE the wait: a
          // shared-memory load issued while the MMAs stream their operands takes hundreds of cycles
          float f[16];
          if (c0 < Np) {
            const float* bh = wblk + a.woff[l] + (c0 >> 5) * lay.Kp[l] * 32 + lay.Kd[l] * 32 + (c0 & 31);
            const float* bl = bh + lay.img_floats[l];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 h4 = *reinterpret_cast<const float4*>(bh + j), l4 = *reinterpret_cast<const float4*>(bl + j);
              f[j] = h4.x + l4.x; f[j + 1] = h4.y + l4.y; f[j + 2] = h4.z + l4.z; f[j + 3] = h4.w + l4.w;
            }
          }
          mbar_wait_warp(&bar_d, ph_d, a.opt);
          ph_d ^= 1;
          umma::tc_fence_after();
          TCB_STAMP(8 + 2 * l);
          if (c0 < Np) {
            uint32_t v[16];
            umma::tmem_ld16(tD + lane_addr + c0, v);
            umma::tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] += __uint_as_float(v[j]);
            tc_act16(a.act[l], f);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(f[j]);
            umma::tmem_st16(tmem + a.c_zs[l + 1] + lane_addr + c0, v);
            if (l < L - 2) {
              uint32_t lo[16];
              tc_split16(f, v, lo);
              umma::tmem_st16(tAhi + lane_addr + c0, v);
              umma::tmem_st16(tAlo + lane_addr + c0, lo);
            }
          }
          umma::tmem_wait_st();
          if (l < L - 2) arrive(&bar_fg);  // the next recomputed layer's A operand is in place
          // layer stream_a's image (MMAs done, bias row read by every worker) makes room for stream_b's, first needed by
          // the input-gradient MMAs of the last layer -- at least one recomputed layer away
          if (streaming && l == a.stream_a) {
            worker_sync();
            if (tid == 0) {
              umma::fence_async_smem();
              load_layer(a.stream_b, &bar_ring);
            }
          }
        }
      }

      TCB_STAMP(1);
      // ---- 2. cotangent of the MLP output: chunk of G_L ----
      float g[16];
      {
        const int Np = lay.Np[L - 1];
#pragma unroll
        for (int j = 0; j < 16; ++j) g[j] = 0.f;
        if (valid && c0 < Np) {
          const float* gp = a.gout_ptr + (size_t)(NODE ? (k0 + row) : d) * a.gout_ld + c0;
          if ((a.gout_ld & 3) == 0 && c0 + 16 <= dout) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const float4 t = *reinterpret_cast<const float4*>(gp + j);
              g[j] = t.x; g[j + 1] = t.y; g[j + 2] = t.z; g[j + 3] = t.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < dout) g[j] = gp[j];
          }
          if (!NODE && aggr == NGPDE_AGGR_MEAN) {
#pragma unroll
            for (int j = 0; j < 16; ++j) g[j] *= rdeg;
          }
          if (EXT && NODE && a.yact != nullptr) {  // an activation behind the MLP (GCNConv): dP = dy * act'(y)
            const float* yp = a.yact + (size_t)(k0 + row) * dout + c0;
            if ((dout & 3) == 0 && c0 + 16 <= dout && (reinterpret_cast<uintptr_t>(a.yact) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 t = *reinterpret_cast<const float4*>(yp + j);
                g[j] *= act_grad_y(a.yact_kind, t.x); g[j + 1] *= act_grad_y(a.yact_kind, t.y);
                g[j + 2] *= act_grad_y(a.yact_kind, t.z); g[j + 3] *= act_grad_y(a.yact_kind, t.w);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < dout) g[j] *= act_grad_y(a.yact_kind, yp[j]);
            }
          }
        }
      }

      // bias gradient = column sums of G: the butterfly is spread over the layer body (see tcb_colsum_a)
      float cw[8];
      tcb_colsum_a(g, lane, cw);
      TCB_STAMP(2);
      // ---- 3. back through the layers.  Per layer: G -> TMEM A operand, then warp 2 (idle while rows 0..63 are staged)
      // issues the input-gradient MMAs; rows 0..63 of G and Z are staged and warp 3 issues their weight-gradient MMAs;
      // rows 64..127 are staged as soon as that batch has drained the buffer, warp 1 issues the second batch; the next G
      // is formed from dZ while it runs and the dW^T block is collected at the top of the following layer.  (The layer
      // loop is deliberately NOT unrolled: the unrolled kernel spent its time in instruction-cache misses.)
      auto collect_dw = [&](int l, uint32_t tDwl) {
        float add[NACC];
#pragma unroll
        for (int j = 0; j < NACC; ++j) add[j] = 0.f;
        if (c0 < min(lay.Kd[l], 64)) {
          uint32_t v[16];
          umma::tmem_ld16(tDwl + lane_addr + c0, v);
          umma::tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float up = __shfl_sync(0xffffffffu, __uint_as_float(v[8 + j]), lane & 15);
            add[j] = upper ? up : __uint_as_float(v[j]);
          }
        }
#pragma unroll
        for (int t = 0; t < NACC - 8; ++t) {
          if (64 + 8 * t < lay.Kd[l]) {
            uint32_t w8[8];
            umma::tmem_ld8(tDwl + lane_addr + 64 + 8 * t, w8);
            umma::tmem_wait_ld();
            const uint32_t ev = q == 0 ? w8[0] : (q == 1 ? w8[2] : (q == 2 ? w8[4] : w8[6]));
            const uint32_t od = q == 0 ? w8[1] : (q == 1 ? w8[3] : (q == 2 ? w8[5] : w8[7]));
            const float up = __shfl_sync(0xffffffffu, __uint_as_float(od), lane & 15);
            add[8 + t] = upper ? up : __uint_as_float(ev);
          }
        }
        // register accumulators need compile-time indices
#pragma unroll
        for (int ll = 0; ll < TCB_MAXL; ++ll) {
          if (ll == l) {
#pragma unroll
            for (int j = 0; j < NACC; ++j) dwacc[ll][j] += add[j];
          }
        }
      };
#pragma unroll 1
      for (int l = L - 1; l >= 0; --l) {
        const int Np = lay.Np[l], Kd = lay.Kd[l];
        const bool active = c0 < Np;
        const bool do_dgrad = (l > 0 || a.need_dz0) && !(id0 && l == 0);
        const uint32_t tDl = (l == 0) ? tmem + a.c_d0 : tD;
        // with two dW^T accumulators (dw_alt) a layer's block is collected two layers later, off the critical path; with one
        // it has to be collected before the next layer's weight-gradient batch is issued
        const bool defer = a.dw_alt > 0;
        const int gz = (Kd + 31) >> 5;  // 32-column groups of this layer's staged Z images
        float* st_zhi = st_base;
        float* st_zlo = st_base + gz * ROWS * 32;
        TCB_STAMP(3 + 6 * l);
        // (A) G_{l+1} -> TMEM A operand; the issuer queues the input-gradient MMAs dZ_l = G W_l^T behind the previous layer's
        // weight-gradient batch
        if (do_dgrad) {
          if (active) {
            uint32_t hi[16], lo[16];
            tc_split16(g, hi, lo);
            umma::tmem_st16(tAhi + lane_addr + c0, hi);
            umma::tmem_st16(tAlo + lane_addr + c0, lo);
          }
          umma::tmem_wait_st();
          arrive(&bar_fg);
        }
        // (B) deferred: the dW^T block of layer l + 2 (complete since the previous iteration's (C); its accumulator is the
        // one this layer's weight-gradient batch will overwrite)
        if (defer && l + 2 <= L - 1) collect_dw(l + 2, tDw + (((l + 2) & 1) ? a.dw_alt : 0));
        // (C) the previous layer's weight-gradient batch has drained the staging buffer
        // identity layer 0 (hoisted): nothing is staged and no batch is issued for it, so layer 1's batch is only waited for
        // (and its dW^T block collected) at the end of the tile, under the cotangent scatter
        const bool lazy0 = id0 && l == 0;
        if (l < L - 1 && !lazy0) {
          mbar_wait_warp(&bar_w, ph_w, a.opt);
          ph_w ^= 1;
          umma::tc_fence_after();
          if (!defer) collect_dw(l + 1, tDw);
        }
        float cy[2] = {0.f, 0.f};
        if (!lazy0) tcb_colsum_b(cw, lane, cy);
        TCB_STAMP(4 + 6 * l);
        // (D) stage G_{l+1} and Z_l as MN-major hi/lo images -- the whole tile at once (FULL), or rows 0..63 and then, once
        // that batch has drained the buffer, rows 64..127 -- while the input-gradient MMAs run
#pragma unroll 1
        for (int h = 0; h < (lazy0 ? 0 : NH); ++h) {
          if (h == 1) {
            mbar_wait_warp(&bar_w, ph_w, a.opt);  // the first half has been consumed
            ph_w ^= 1;
          }
          if ((FULL || (lq >> 1) == h) && !(skip_w0 && l == 0)) {
            const int r = row - ROWS * h;
            if (active) stage_chunk<ROWS>(st_ghi, st_glo, r, c0, g);
            // Z_l: the gathered input for l == 0, else the FP32 copy kept in TMEM by the recompute
            for (int cc = c0; cc < Kd; cc += 64) {
              float z[16];
              if (l == 0) {
                if (keep_z0) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) z[j] = DZ[row * (Kd0 + 1) + cc + j];
                } else {
                  const TcChunk ch = NODE ? chunks[cc >> 4] : TcChunk{nullptr, 0};
                  if (NODE && ch.base != nullptr) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) z[j] = 0.f;
                    if (valid) tc_load_chunk<4>(ch, d, 0, z);
                  } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) z[j] = valid ? tc_gather_col<EXT>(cols[cc + j], s, d, p, pg) : 0.f;
                  }
                }
              } else {
                uint32_t v[16];
                umma::tmem_ld16(tmem + a.c_zs[l] + lane_addr + cc, v);
                umma::tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) z[j] = __uint_as_float(v[j]);
              }
              stage_chunk<ROWS>(st_zhi, st_zlo, r, cc, z);
            }
          }
          umma::fence_async_smem();
          if (h == 0) {
            TCB_STAMP(5 + 6 * l);
            // (E) bias gradient: column sums of G over this warp's 32 rows
            {
              const float cs = tcb_colsum_c(cy, lane);
              if (active) {
#pragma unroll
                for (int ll = 0; ll < TCB_MAXL; ++ll)
                  if (ll == l) dbacc[ll] += cs;
              }
            }
            // (F) dZ_l is complete.  Waited for BEFORE the weight-gradient batch is released: the tensor pipe runs its
            // batches in order anyway, and once the shared-memory-bound SS MMAs are running an mbarrier poll from these
            // warps is not served until they end (measured: tools/tcb_phases.py)
            TCB_STAMP(52 + l);
            if (do_dgrad) {
              if (TCB_DIAG && a.dbg != nullptr && blockIdx.x == 0 && dbg_tile < 8) {  // diagnosis: poll vs warp re-convergence
                if (lane == 0) {
                  umma::mbar_spin(&bar_d, ph_d);
                  if (tid == TCB_STAMP_TID) a.dbg[dbg_tile * 64 + 56 + l] = clock64();
                }
                __syncwarp();
              } else {
                mbar_wait_warp(&bar_d, ph_d, a.opt);
              }
              ph_d ^= 1;
            }
            TCB_STAMP(6 + 6 * l);
          }
          // (G) staged images ready: the issuer releases the weight-gradient MMAs dW_l^T += G^T Z
          arrive(&bar_fs);
        }
        // the last layer's image has served its only use: bring stream_a's back (needed again by its own input-gradient
        // MMAs further down and by the next tile's recompute)
        if (streaming && l == a.stream_b && tid == 0) {
          umma::fence_async_smem();
          load_layer(a.stream_a, &bar_ring);
        }
        TCB_STAMP(7 + 6 * l);
        umma::tc_fence_after();
        if (l > 0) {
          // G_l = dZ_l .* act'(Z_l) for this thread's chunk of layer l-1's output
#pragma unroll
          for (int j = 0; j < 16; ++j) g[j] = 0.f;
          if (c0 < lay.Np[l - 1]) {  // warp-uniform: tcgen05.ld is .sync.aligned (never put `valid` in this condition)
            uint32_t v[16], zz[16];
            umma::tmem_ld16(tDl + lane_addr + c0, v);
            umma::tmem_ld16(tmem + a.c_zs[l] + lane_addr + c0, zz);
            umma::tmem_wait_ld();
            const int act = a.act[l - 1];
            if (act == NGPDE_ACT_TANH) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float y = __uint_as_float(zz[j]);
                g[j] = __uint_as_float(v[j]) * fmaf(-y, y, 1.f);
              }
            } else if (act == NGPDE_ACT_IDENTITY) {
#pragma unroll
              for (int j = 0; j < 16; ++j) g[j] = __uint_as_float(v[j]);
            } else {
#pragma unroll 1
              for (int j = 0; j < 16; ++j) zz[j] = __float_as_uint(act_grad_y(act, __uint_as_float(zz[j])));
#pragma unroll
              for (int j = 0; j < 16; ++j) g[j] = __uint_as_float(v[j]) * __uint_as_float(zz[j]);
            }
            if (!valid) {
#pragma unroll
              for (int j = 0; j < 16; ++j) g[j] = 0.f;
            }
          }
          if (!(id0 && l == 1)) tcb_colsum_a(g, lane, cw);  // (the identity layer has no bias)
        } else {
          if (a.need_dz0) {
            // ---- 4. dZ_0 back to its sources ----
            for (int cc = c0; cc < Kd0; cc += 64) {
              uint32_t v[16];
              if (id0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(g[j]);
              } else {
                umma::tmem_ld16(tDl + lane_addr + cc, v);
                umma::tmem_wait_ld();
              }
              if (NODE) {
                const TcChunk dch = dchunks[cc >> 4];
                if (valid && dch.base != nullptr) {
                  float4* o = reinterpret_cast<float4*>(const_cast<float*>(dch.base) + (size_t)(k0 + row) * dch.ld);
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                       __uint_as_float(v[4 * j + 3]));
                } else if (valid) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const TcDst t = dcols[cc + j];
                    if (t.base != nullptr) t.base[(size_t)(k0 + row) * t.ld] = __uint_as_float(v[j]);
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) DZ[row * (Kd0 + 1) + cc + j] = __uint_as_float(v[j]);
                if (direct_src && valid) {  // desrc[k0 + row][cc .. cc + 16): this thread's registers
                  float4* o = reinterpret_cast<float4*>(a.desrc + (size_t)(k0 + row) * a.src_w + cc);
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    o[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                       __uint_as_float(v[4 * j + 3]));
                }
              }
            }
          }
          // layer 1's block (deferred); layer 0's closes the tile, after the scatter below (its MMAs are still running)
          if (a.dw_alt > 0 && L > 1 && !id0) collect_dw(1, tDw + a.dw_alt);
        }
        umma::tc_fence_before();
      }

      if (!NODE && a.need_dz0) {
        worker_sync();
        const int dx = a.dx, ldz = Kd0 + 1;
        // source side: one row per edge (the x columns that have a source-side use), reduced later over the src-sorted transpose
        const int sw = a.src_w, sc0 = a.src_c0;
        auto src_cot = [&](int e, int cc) {
          const int c = sc0 + cc;
          float v = 0.f;
          for (int si = 0; si < a.n_segs; ++si) {
            const Seg sg = a.segs[si];
            if (sg.arr != ARR_X || c < sg.col || c >= sg.col + sg.width) continue;
            const float cf = coef_src(sg.kind);
            if (cf != 0.f) v = fmaf(cf, DZ[e * ldz + sg.row + c - sg.col], v);
          }
          a.desrc[(size_t)(k0 + e) * sw + cc] = v;
        };
        // Wide rows (hoisted first layer: 64 columns per edge): the per-element walk over the segment table was the most
        // expensive phase of the tile (11k of 37k cycles at C5).  A lane owns a column, decodes its (at most two) contributing
        // segments ONCE, and then walks the edges: a load, one or two multiplies and a coalesced store per element.
        auto decode = [&](int c, bool src_side, int& z0, float& f0, int& z1, float& f1) {
          int n = 0;
          z0 = z1 = 0;
          f0 = f1 = 0.f;
          for (int si = 0; si < a.n_segs; ++si) {
            const Seg sg = a.segs[si];
            if (sg.arr != ARR_X || c < sg.col || c >= sg.col + sg.width) continue;
            const float cf = src_side ? coef_src(sg.kind) : coef_dst(sg.kind);
            if (cf == 0.f) continue;
            if (n == 0) { z0 = sg.row + c - sg.col; f0 = cf; } else { z1 = sg.row + c - sg.col; f1 = cf; }
            ++n;
          }
          return n;
        };
        if (direct_src) {
          // already stored from the registers that held dZ_0 (step 4)
        } else if (EXT && sw >= 32) {
          for (int cc = tid & 31; cc < sw; cc += 32) {
            int z0, z1;
            float f0, f1;
            const int n = decode(sc0 + cc, true, z0, f0, z1, f1);
            if (n > 2) {
              for (int e = tid >> 5; e < ne; e += NW / 32) src_cot(e, cc);
              continue;
            }
            for (int e = tid >> 5; e < ne; e += NW / 32) {
              float v = n > 0 ? fmaf(f0, DZ[e * ldz + z0], 0.f) : 0.f;
              if (n > 1) v = fmaf(f1, DZ[e * ldz + z1], v);
              a.desrc[(size_t)(k0 + e) * sw + cc] = v;
            }
          }
        } else {
          for (int item = tid; item < ne * sw; item += NW) {
            const int e = item / sw;
            src_cot(e, item - e * sw);
          }
        }
        TCB_STAMP(28);
        // destination side: sequential over the row's edges, carried across tiles through dxdst itself
        if (EXT && a.has_dst_side && a.dst_w >= 32) {  // wide: a warp per destination row, lanes along the columns that have a destination side
          for (int jj = tid >> 5; jj < n1 - n0; jj += NW / 32) {
            const int j = n0 + jj;
            const int r0 = a.tg.rowptr[j], r1 = a.tg.rowptr[j + 1];
            const int lo = max(r0, k0), hi = min(r1, k0 + ne);
            if (lo >= hi) continue;
            for (int c = a.dst_c0 + (tid & 31); c < a.dst_c0 + a.dst_w; c += 32) {
              int z0, z1;
              float f0, f1;
              const int n = decode(c, false, z0, f0, z1, f1);
              float accv = (lo == r0) ? 0.f : a.dxdst[(size_t)j * dx + c];
              if (n <= 2) {
                if (n > 0) for (int e = lo; e < hi; ++e) accv = fmaf(f0, DZ[(e - k0) * ldz + z0], accv);
                if (n > 1) for (int e = lo; e < hi; ++e) accv = fmaf(f1, DZ[(e - k0) * ldz + z1], accv);
              } else {
                for (int si = 0; si < a.n_segs; ++si) {
                  const Seg sg = a.segs[si];
                  if (sg.arr != ARR_X || c < sg.col || c >= sg.col + sg.width) continue;
                  const float cf = coef_dst(sg.kind);
                  if (cf == 0.f) continue;
                  const float* gr = DZ + sg.row + c - sg.col;
                  for (int e = lo; e < hi; ++e) accv = fmaf(cf, gr[(e - k0) * ldz], accv);
                }
              }
              a.dxdst[(size_t)j * dx + c] = accv;
            }
          }
        } else if (a.has_dst_side) {
          for (int item = tid; item < (n1 - n0) * dx; item += NW) {
            const int jj = item / dx, c = item - jj * dx;
            const int j = n0 + jj;
            const int r0 = a.tg.rowptr[j], r1 = a.tg.rowptr[j + 1];
            const int lo = max(r0, k0), hi = min(r1, k0 + ne);
            if (lo >= hi) continue;
            float accv = (lo == r0) ? 0.f : a.dxdst[(size_t)j * dx + c];
            for (int si = 0; si < a.n_segs; ++si) {
              const Seg sg = a.segs[si];
              if (sg.arr != ARR_X || c < sg.col || c >= sg.col + sg.width) continue;
              const float cf = coef_dst(sg.kind);
              if (cf == 0.f) continue;
              const float* gr = DZ + sg.row + c - sg.col;
              for (int e = lo; e < hi; ++e) accv = fmaf(cf, gr[(e - k0) * ldz], accv);
            }
            a.dxdst[(size_t)j * dx + c] = accv;
          }
        }
      }
      TCB_STAMP(29);
      // layer 0's dW^T block: its weight-gradient batch ran under the scatter above
      mbar_wait_warp(&bar_w, ph_w, a.opt);  // id0: layer 1's last batch (layer 0 issued none), else layer 0's
      TCB_STAMP(30);
      ph_w ^= 1;
      umma::tc_fence_after();
      if (id0) collect_dw(1, a.dw_alt > 0 ? tDw + a.dw_alt : tDw);
      else if (!skip_w0) collect_dw(0, tmem + a.c_dw0);
      umma::tc_fence_before();
      worker_sync();
      TCB_STAMP(27);
      ++dbg_tile;
    }
  }

  // ---- this CTA's parameter-gradient partial ----
  {
    float* dWp = a.dparams_partial + (size_t)blockIdx.x * a.n_params;
    const int n = 16 * lq + (lane & 15);
#pragma unroll
    for (int l = 0; l < TCB_MAXL; ++l) {
      if (l >= L) continue;
      const int K = lay.K[l], N = lay.N[l];
      if (n >= N) continue;
      auto put = [&](int k, float v) {
        if (k < K) dWp[a.w_off[l] + (size_t)(l == 0 ? tc_orig_row(lay, k) : k) * N + n] = v;
      };
#pragma unroll
      for (int j = 0; j < 8; ++j) put(c0 + (upper ? 8 : 0) + j, dwacc[l][j]);
#pragma unroll
      for (int t = 0; t < NACC - 8; ++t) put(64 + 8 * t + 2 * q + (upper ? 1 : 0), dwacc[l][8 + t]);
    }
    // bias gradients: the four row quarters of a column are summed in fixed order through shared memory
    float* dbx = reinterpret_cast<float*>(smem + a.off_stage);  // [L][4 quarters][64 columns]; the staging buffer is free now
    worker_sync();
    if ((lane & 1) == 0) {
      const int col = c0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
#pragma unroll
      for (int l = 0; l < TCB_MAXL; ++l)
        if (l < L) dbx[(l * 4 + lq) * 64 + col] = dbacc[l];
    }
    worker_sync();
    for (int item = tid; item < L * 64; item += NW) {
      const int l = item >> 6, col = item & 63;
      if (col < lay.N[l] && a.b_off[l] >= 0) {
        const float* pq = dbx + l * 256 + col;
        dWp[a.b_off[l] + col] = ((pq[0] + pq[64]) + pq[128]) + pq[192];
      }
    }
  }
  }  // workers
  umma::tc_fence_before();
  __syncthreads();
  if (tid < 32) umma::tmem_dealloc(tmem, a.tmem_cols);
}

}  // namespace ngpde
