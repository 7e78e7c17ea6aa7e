// Do MMA batches of two issuing warps complete in issue order, and when does each commit's mbarrier fire?
// warp 0: 24 TS MMAs (dgrad form) -> commit bar_a.   warp 1: (after `delay` cycles) 48 SS MMAs (wgrad form) -> commit bar_b.
// warp 2 polls bar_a, warp 3 polls bar_b; all stamp clock64.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../neuralgraphpde.jl_b200/csrc/ngpde_umma.cuh"
using namespace ngpde::umma;

__global__ void __launch_bounds__(512) order_kernel(int delay, int second_is_ts, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar_a, bar_b;
  const int tid = threadIdx.x, warp = uniform_i32(threadIdx.x >> 5);
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { mbar_init(&bar_a, 1); mbar_init(&bar_b, 1); fence_mbar_init(); }
  for (int i = tid; i < 160 * 1024 / 4; i += 512) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 255);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = uniform_u32(tmem_slot);
  const uint32_t sb = smem_u32(smem);
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint32_t idesc = make_idesc(128, 64, 0, 0);
      const uint64_t db = make_sdesc(sb, 0, 512, 1);
      const uint32_t g16 = 8u * 72;
      out[0] = clock64() - t0;
      for (int pass = 0; pass < 3; ++pass)
#pragma unroll 4
        for (int k = 0; k < 8; ++k)
          mma_tf32_ts(tmem, tmem + 256 + k * 8, db + (uint64_t)((k >> 2) * g16 + (k & 3) * 2u), idesc, (pass | k) != 0);
      mma_commit(&bar_a);
      out[1] = clock64() - t0;
      if (delay < 0) {  // same thread issues the second batch right behind
        out[2] = clock64() - t0;
        const uint32_t idesc2 = make_idesc(64, 64, 1, 1);
        const uint32_t lbo = 128u * 128;
        const uint64_t da = make_sdesc(sb + 65536, lbo, 512, 1), dbb = make_sdesc(sb + 98304, lbo, 512, 1);
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll 8
          for (int k = 0; k < 16; ++k) mma_tf32_ss(tmem + 128, da + (uint64_t)(k * 64), dbb + (uint64_t)(k * 64), idesc2, (pass | k) != 0);
        mma_commit(&bar_b);
        out[3] = clock64() - t0;
      }
    }
    __syncwarp();
  } else if (warp == 1 && delay >= 0) {
    if (elect_one_sync()) {
      while (clock64() - t0 < delay) {}
      out[2] = clock64() - t0;
      if (second_is_ts) {
        const uint32_t idesc = make_idesc(128, 64, 0, 1);
        const uint64_t db = make_sdesc(sb + 65536, 128u * 72, 512, 1);
        for (int pass = 0; pass < 6; ++pass)
#pragma unroll 4
          for (int k = 0; k < 8; ++k) mma_tf32_ts(tmem + 128, tmem + 384 + k * 8, db + (uint64_t)(k * 64), idesc, (pass | k) != 0);
      } else {
        const uint32_t idesc = make_idesc(64, 64, 1, 1);
        const uint32_t lbo = 128u * 128;
        const uint64_t da = make_sdesc(sb + 65536, lbo, 512, 1), dbb = make_sdesc(sb + 98304, lbo, 512, 1);
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll 8
          for (int k = 0; k < 16; ++k) mma_tf32_ss(tmem + 128, da + (uint64_t)(k * 64), dbb + (uint64_t)(k * 64), idesc, (pass | k) != 0);
      }
      mma_commit(&bar_b);
      out[3] = clock64() - t0;
    }
    __syncwarp();
  } else if (warp == 2) {
    if ((tid & 31) == 0) { mbar_spin(&bar_a, 0); out[4] = clock64() - t0; }
    __syncwarp();
  } else if (warp == 3) {
    if ((tid & 31) == 0) { mbar_spin(&bar_b, 0); out[5] = clock64() - t0; }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  long long h[8];
  cudaFuncSetAttribute(order_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int second_ts : {0, 1})
    for (int delay : {-1, 0, 300, 800, 1500, 3000}) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaMemset(d, 0, 64);
        order_kernel<<<1, 512, 180 * 1024>>>(delay, second_ts, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        if (rep == 1)
          printf("second=%s delay %4d: A issue %lld..%lld  B issue %lld..%lld | bar_a fires %lld  bar_b fires %lld  (%s)\n",
                 second_ts ? "TS x48" : "SS M=64 x48", delay, h[0], h[1], h[2], h[3], h[4], h[5], cudaGetErrorString(e));
      }
    }
  return 0;
}
