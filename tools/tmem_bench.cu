// Timing probe for the tensor-core message-passing kernels (run on a B200): measures, in SM cycles (clock64),
//   * tcgen05.ld / tcgen05.st throughput for 4, 8 and 16 warps and the x8/x16/x32/x64 shapes,
//   * the duration of the MMA batches the kernels issue (M = 128 TS N = 64; M = 64 SS N = 72; M = 128 SS N = 72),
//   * tcgen05.ld throughput while an MMA batch is running (TMEM port contention).
// The operands are whatever happens to be in shared memory / TMEM: only time is measured.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/tmem_bench tools/tmem_bench.cu && tools/bin/tmem_bench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../neuralgraphpde.jl_b200/csrc/ngpde_umma.cuh"

using namespace ngpde::umma;

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
        "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
        "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// mode 0: ld x8, 1: ld x16, 2: ld x32, 3: st x16, 4: st x8
__global__ void __launch_bounds__(512) ldst_kernel(int mode, int nwarps, int reps, long long* out) {
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int r = 0; r < reps; ++r) {
      if (mode == 0) {
        uint32_t v[8];
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
          tmem_ld8(base + c, v);
          tmem_wait_ld();
          acc += v[0] ^ v[7];
        }
      } else if (mode == 1) {
        uint32_t v[16];
#pragma unroll
        for (int c = 0; c < 64; c += 16) {
          tmem_ld16(base + c, v);
          tmem_wait_ld();
          acc += v[0] ^ v[15];
        }
      } else if (mode == 2) {
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          tmem_ld32(base + c, v);
          tmem_wait_ld();
          acc += v[0] ^ v[31];
        }
      } else if (mode == 5) {  // two x16 loads in flight before the wait
        uint32_t v[16], w[16];
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          tmem_ld16(base + c, v);
          tmem_ld16(base + c + 16, w);
          tmem_wait_ld();
          acc += v[0] ^ w[15];
        }
      } else if (mode == 3) {
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = acc + j;
#pragma unroll
        for (int c = 0; c < 64; c += 16) tmem_st16(base + c, v);
        tmem_wait_st();
      } else {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = acc + j;
#pragma unroll
        for (int c = 0; c < 64; c += 8) tmem_st8(base + c, v);
        tmem_wait_st();
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) out[0] = t1 - t0;
  if (acc == 0x12345678u) out[1] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// MMA batches.  kind 0: TS M=128 N=n K-steps=ks (A in TMEM, B smem MN-major image); kind 1: SS M=64; kind 2: SS M=128.
// ld_warps > 0: warps 4.. run x16 loads over columns 256.. while the MMAs execute; their time is reported too.
__global__ void __launch_bounds__(512) mma_kernel(int kind, int n, int nmma, int reps, int ld_warps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 160 * 1024 / 4; i += 512) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 255);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t sb = smem_u32(smem);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0, t2 = 0;
  __syncthreads();
  t0 = clock64();
  if (tid == 0) {
    uint32_t ph = 0;
    for (int r = 0; r < reps; ++r) {
      if (kind == 0) {
        const uint32_t idesc = make_idesc(128, n, 0, 1);
        const uint64_t db = make_sdesc(sb, 128u * 72, 512, 1);
        for (int i = 0; i < nmma; ++i) mma_tf32_ts(tmem, tmem + 64 + (i % 9) * 8, db + (uint64_t)((i % 9) * 64), idesc, i > 0);
      } else if (kind == 3 || kind == 4) {  // 3: alternate between two accumulators; 4: same but unrolled, 8 K-steps
        const uint32_t idesc = make_idesc(128, n, 0, 1);
        const uint64_t db = make_sdesc(sb, 128u * 72, 512, 1);
        if (kind == 3) {
          for (int i = 0; i < nmma; ++i)
            mma_tf32_ts(tmem + (i & 1) * 128, tmem + 256 + (i & 7) * 8, db + (uint64_t)((i & 7) * 64), idesc, i > 1);
        } else {
          for (int i0 = 0; i0 < nmma; i0 += 8) {
#pragma unroll
            for (int i = 0; i < 8; ++i) mma_tf32_ts(tmem, tmem + 256 + i * 8, db + (uint64_t)(i * 64), idesc, (i0 | i) > 0);
          }
        }
      } else if (kind == 6) {  // dgrad-style: B read K-major from the MN-major image (SBO = 512, LBO = 0), N = n rows
        const uint32_t idesc = make_idesc(128, n, 0, 0);
        const uint64_t db = make_sdesc(sb, 0, 512, 1);
        const uint32_t g16 = 8u * 72;
        for (int i0 = 0; i0 < nmma; i0 += 8) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
            mma_tf32_ts(tmem, tmem + 256 + i * 8, db + (uint64_t)((i >> 2) * g16 + (i & 3) * 2u), idesc, (i0 | i) > 0);
        }
      } else if (kind == 5) {
        const uint32_t idesc = make_idesc(128, n, 0, 1);
        const uint64_t db = make_sdesc(sb, 128u * 72, 512, 1);
        for (int i0 = 0; i0 < nmma; i0 += 8) {
#pragma unroll
          for (int i = 0; i < 8; ++i) mma_tf32_ts(tmem, tmem + 256 + i * 8, db + (uint64_t)(i * 64), idesc, (i0 | i) > 0);
        }
      } else {
        const uint32_t idesc = make_idesc(kind == 1 ? 64 : 128, n, 1, 1);
        const uint64_t da = make_sdesc(sb, 128u * 64, 512, 1), db = make_sdesc(sb + 65536, 128u * 64, 512, 1);
        for (int i = 0; i < nmma; ++i) mma_tf32_ss(tmem, da + (uint64_t)((i % 8) * 64), db + (uint64_t)((i % 8) * 64), idesc, i > 0);
      }
      t2 = clock64();
      mma_commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
    }
    t1 = clock64();
    out[0] = t1 - t0;   // all batches, issue + completion
    out[2] = t2 - t0;   // issue only (last batch's issue end)
  } else if (kind == 5 && tid == 32) {  // second issuer, own accumulator and barrier
    __shared__ __align__(8) uint64_t bar2;
    mbar_init(&bar2, 1);
    fence_mbar_init();
    uint32_t ph = 0;
    const uint32_t idesc = make_idesc(128, n, 0, 1);
    const uint64_t db = make_sdesc(sb + 32768, 128u * 72, 512, 1);
    for (int r = 0; r < reps; ++r) {
      for (int i0 = 0; i0 < nmma; i0 += 8) {
#pragma unroll
        for (int i = 0; i < 8; ++i) mma_tf32_ts(tmem + 128, tmem + 384 + i * 8, db + (uint64_t)(i * 64), idesc, (i0 | i) > 0);
      }
      mma_commit(&bar2);
      mbar_wait(&bar2, ph);
      ph ^= 1;
    }
  } else if (warp >= 4 && warp < 4 + ld_warps) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256 + ((warp - 4) >> 2) * 64;
    const int lreps = reps * nmma / 4;
    for (int r = 0; r < lreps; ++r) {
      uint32_t v[16];
#pragma unroll
      for (int c = 0; c < 64; c += 16) {
        tmem_ld16(base + c, v);
        tmem_wait_ld();
        acc += v[0] ^ v[15];
      }
    }
    t1 = clock64();
    if ((tid & 31) == 0 && warp == 4) out[1] = t1 - t0;
  }
  if (acc == 0x12345678u) out[3] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  long long h[4];
  const char* names[] = {"ld x8", "ld x16", "ld x32", "st x16", "st x8", "ld 2x16"};
  for (int mode : {0, 1, 2, 5, 3, 4}) {
    for (int nw : {4, 8, 16}) {
      const int reps = 64;
      cudaMemset(d, 0, 64);
      ldst_kernel<<<1, 512>>>(mode, nw, reps, d);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
      const double bytes = (double)nw * 32 * 64 * 4 * reps;
      printf("%-8s warps %2d: %8lld cycles, %6.1f B/cycle/SM  (%s)\n", names[mode], nw, h[0], bytes / h[0],
             cudaGetErrorString(e));
    }
  }
  cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Case { int kind, n, nmma; const char* what; };
  const Case cases[] = {{0, 64, 27, "TS M=128 N=64 (forward layer, 27 MMAs)"}, {0, 64, 9, "TS M=128 N=64, 9 MMAs"},
                        {0, 72, 24, "TS M=128 N=72 (dgrad, 24 MMAs)"},
                        {1, 72, 24, "SS M=64 N=72 (wgrad half, 24 MMAs)"}, {2, 72, 24, "SS M=128 N=72, 24 MMAs"},
                        {2, 64, 24, "SS M=128 N=64, 24 MMAs"}, {1, 64, 24, "SS M=64 N=64, 24 MMAs"},
                        {2, 144, 24, "SS M=128 N=144, 24 MMAs"}, {2, 256, 24, "SS M=128 N=256, 24 MMAs"},
                        {4, 64, 24, "TS M=128 N=64 unrolled issue, 24 MMAs"}, {4, 128, 24, "TS M=128 N=128 unrolled, 24 MMAs"},
                        {4, 32, 24, "TS M=128 N=32 unrolled, 24 MMAs"}, {4, 16, 24, "TS M=128 N=16 unrolled, 24 MMAs"},
                        {4, 96, 24, "TS M=128 N=96 unrolled, 24 MMAs"},
                        {3, 64, 24, "TS M=128 N=64 two accumulators, 24 MMAs"},
                        {6, 64, 24, "TS M=128 N=64 B K-major (dgrad), 24 MMAs"},
                        {6, 16, 24, "TS M=128 N=16 B K-major (dgrad L0), 24 MMAs"},
                        {5, 64, 24, "TS M=128 N=64 two issuing warps, 24 MMAs each"}};
  for (const Case& c : cases) {
    for (int ldw : {0, 12}) {
      const int reps = 16;
      cudaMemset(d, 0, 64);
      mma_kernel<<<1, 512, 180 * 1024>>>(c.kind, c.n, c.nmma, reps, ldw, d);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
      printf("%-48s ld warps %2d: %7.1f cycles/batch (%5.1f per MMA; issue-only clock to last issue %7.1f/batch)", c.what, ldw,
             (double)h[0] / reps, (double)h[0] / reps / c.nmma, (double)h[2] / reps);
      if (ldw) printf("; concurrent ld: %6.1f B/cycle/SM", (double)ldw * 32 * 64 * 4 * (reps * c.nmma / 4) / (double)h[1]);
      printf("  (%s)\n", cudaGetErrorString(e));
    }
  }
  return 0;
}
