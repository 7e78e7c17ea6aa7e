// tcgen05 / TMEM / mbarrier primitives for the tensor-core message-passing kernels (sm_100a only, inline PTX).
//
// Conventions pinned on hardware by tools/umma_probe.cu:
//   * accumulators and TMEM-resident A operands of an M = 128 MMA: TMEM lane = row (edge), column = n (resp. k);
//   * shared-memory operands are SWIZZLE_128B blocks of 128-byte rows: element (r, c), c in [0, 32) floats, of group g
//     lives at float offset  g*rows*32 + r*32 + ((c/4) ^ (r & 7))*4 + c%4  (`sw128_offset`); a group holds 32 floats of
//     the operand's contiguous dimension, groups are LBO bytes apart, 8-row bundles SBO = 1024 bytes apart;
//   * the same image of a weight W[k][n] serves  B = W  (MN-major, forward)  and  B = W^T  (K-major, input gradient).
// FP32 accuracy comes from the 3xTF32 split a = hi + lo + r with hi = a ROUNDED to TF32 (10 explicit mantissa bits,
// |a - hi| <= 2^-11 |a|) and lo = (a - hi) rounded to TF32 (|r| <= 2^-22 |a|):
//   a*b ~= lo_a*hi_b + hi_a*lo_b + hi_a*hi_b   (dropped terms <= 3 * 2^-22 |ab|, unbiased; measured: a truncating split is
//   4x worse and biased).  The small cross terms are issued first: the tensor core truncates its FP32 accumulator at every
//   step, so they are added while the accumulator is still small.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ngpde {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// round to nearest TF32 (ties away from zero): add half an ulp of the 13 dropped bits, then clear them
__host__ __device__ __forceinline__ float tf32_hi(float x) {
#ifdef __CUDA_ARCH__
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
#else
  union { float f; uint32_t u; } v;
  v.f = x;
  v.u = (v.u + 0x1000u) & 0xFFFFE000u;
  return v.f;
#endif
}

// low part of the split.  The tensor core ignores the 13 low mantissa bits of a TF32 operand, so leaving x - hi unrounded
// truncates lo instead of rounding it: an error of at most 2^-11 |lo| <= 2^-22 |x|, the size of the dropped lo*lo term,
// for two integer instructions less per element in every epilogue (set to 1 to round).
#ifndef NGPDE_TF32_ROUND_LO
#define NGPDE_TF32_ROUND_LO 0
#endif
__host__ __device__ __forceinline__ float tf32_lo(float x, float hi) {
#if NGPDE_TF32_ROUND_LO
  return tf32_hi(x - hi);
#else
  return x - hi;
#endif
}

__host__ __device__ __forceinline__ uint32_t sw128_offset(int group, int rows, int r, int c) {
  return (uint32_t)(group * rows * 32 + r * 32 + ((((c >> 2) ^ (r & 7))) << 2) + (c & 3));
}

// SWIZZLE_128B_BASE32B (the only swizzled layout tcgen05 accepts for MN-major 32-bit operands): 32-byte chunks of a
// 128-byte row are XOR-permuted with (row & 3); atoms are 4 rows = 512 bytes.
__host__ __device__ __forceinline__ uint32_t sw128b32_offset(int group, int rows, int r, int c) {
  return (uint32_t)(group * rows * 32 + r * 32 + ((((c >> 3) ^ (r & 3))) << 3) + (c & 7));
}

// instruction descriptor, kind::tf32, FP32 accumulate (cute::UMMA::InstrDescriptor bit layout)
__host__ __device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// shared-memory matrix descriptor, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout, version 1)
// layout_type: 0 = no swizzle, 1 = SWIZZLE_128B_BASE32B, 2 = SWIZZLE_128B, 4 = SWIZZLE_64B, 6 = SWIZZLE_32B
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}
__device__ __forceinline__ uint64_t make_sdesc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return make_sdesc(smem_addr, lbo_bytes, sbo_bytes, 2);
}

// ---- TMEM allocation (one warp, all lanes) ----
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- fences ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// busy poll (test_wait returns at once; try_wait may suspend the thread and wake it late)
__device__ __forceinline__ void mbar_spin(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "SPIN_LOOP:\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra SPIN_DONE;\n\t"
      "bra SPIN_LOOP;\n\t"
      "SPIN_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---- warp-uniform helpers ----
// tcgen05.mma takes its operands from UNIFORM registers.  Issued under a divergent `if (lane == 0)` the compiler wraps every
// MMA in an ELECT / R2UR.BROADCAST x6 / BRA.U.ANY "waterfall" (~65 cycles per MMA, measured); issued under elect.sync with
// operands the compiler can prove warp-uniform it is a plain back-to-back UTCHMMA stream.
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// tells the compiler that `x` is the same in every lane of the (converged) warp
__device__ __forceinline__ uint32_t uniform_u32(uint32_t x) { return __shfl_sync(0xffffffffu, x, 0); }
__device__ __forceinline__ int uniform_i32(int x) { return __shfl_sync(0xffffffffu, x, 0); }

// ---- MMA issue (one thread) ----
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, int accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, int accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on `bar` once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM <-> registers: each thread of warp w touches lane 32*(w%4) + laneid, `n` consecutive 32-bit columns ----
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace umma
}  // namespace ngpde
