// Tensor-pipe rate of the MMA forms the message-passing kernels issue, measured the way the kernels issue them (one elected
// lane under elect.sync, warp-uniform operands), alone and with the other 15 warps streaming 128-bit shared-memory stores
// (the weight-gradient staging) at the same time.  Operands are whatever is in shared memory / TMEM: only time is measured.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/mma_rate tools/mma_rate.cu && tools/bin/mma_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../neuralgraphpde.jl_b200/csrc/ngpde_umma.cuh"
using namespace ngpde::umma;

// form 0: TS  M=128, B MN-major image [K rows][N] (forward / recompute)         N = n, K-steps = ks per pass, 3 passes
// form 1: TS  M=128, B = same image read K-major (input gradient)               N = n
// form 2: SS  M=64,  A MN-major [rows][64], B MN-major [rows][n] (weight grad)  K = rows staged (ks = rows / 8), 3 passes
// form 3: SS  M=128, same operands
__global__ void __launch_bounds__(512) rate_kernel(int form, int n, int ks, int reps, int store_warps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  __shared__ int stop;
  const int tid = threadIdx.x, warp = uniform_i32(threadIdx.x >> 5);
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    stop = 0;
  }
  for (int i = tid; i < 128 * 1024 / 4; i += 512) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 255);
  for (int i = tid; i < 8 * 1024; i += 512) reinterpret_cast<int*>(smem + 128 * 1024)[i] = (i * 7) & 63;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = uniform_u32(tmem_slot);
  const uint32_t sb = smem_u32(smem);
  __syncthreads();
  if (warp == 0) {
    long long t0 = 0, t1 = 0;
    if (elect_one_sync()) {
      uint32_t ph = 0;
      t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        if (form == 0) {
          const uint32_t idesc = make_idesc(128, n, 0, 1);
          const uint64_t db = make_sdesc(sb, 128u * 72, 512, 1);
          for (int pass = 0; pass < 3; ++pass)
#pragma unroll 4
            for (int k = 0; k < ks; ++k) mma_tf32_ts(tmem, tmem + 256 + k * 8, db + (uint64_t)(k * 64), idesc, (pass | k) != 0);
        } else if (form == 1) {
          const uint32_t idesc = make_idesc(128, n, 0, 0);
          const uint64_t db = make_sdesc(sb, 0, 512, 1);
          const uint32_t g16 = 8u * 72;
          for (int pass = 0; pass < 3; ++pass)
#pragma unroll 4
            for (int k = 0; k < ks; ++k)
              mma_tf32_ts(tmem, tmem + 256 + k * 8, db + (uint64_t)((k >> 2) * g16 + (k & 3) * 2u), idesc, (pass | k) != 0);
        } else {
          const uint32_t idesc = make_idesc(form == 2 ? 64 : 128, n, 1, 1);
          const uint32_t lbo = 128u * 8 * ks;
          const uint64_t da = make_sdesc(sb, lbo, 512, 1), db = make_sdesc(sb + 65536, lbo, 512, 1);
          for (int pass = 0; pass < 3; ++pass)
#pragma unroll 8
            for (int k = 0; k < ks; ++k) mma_tf32_ss(tmem, da + (uint64_t)(k * 64), db + (uint64_t)(k * 64), idesc, (pass | k) != 0);
        }
        mma_commit(&bar);
        mbar_wait(&bar, ph);
        ph ^= 1;
      }
      t1 = clock64();
      out[0] = t1 - t0;
      stop = 1;
    }
    __syncwarp();
  } else if (store_warps < 0 && warp <= -store_warps / 100) {
    const int mode = (-store_warps) % 100;
    if (mode == 1) {
      // what a waiting warp does in the kernels: lane 0 polls an mbarrier with test_wait, the others park on __syncwarp
      long long cnt = 0;
      if ((tid & 31) == 0) {
        while (*(volatile int*)&stop == 0) {
          uint32_t ok;
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                       : "=r"(ok) : "r"(smem_u32(&bar)), "r"(1u) : "memory");
          cnt += ok + 1;
        }
      }
      __syncwarp();
      if (tid == 32) out[1] = cnt;
    } else if (mode == 3) {
      // latency of a dependent shared-memory load chain while the MMAs run (lane 0 of every warp)
      long long cnt = 0;
      if ((tid & 31) == 0) {
        volatile int* p = reinterpret_cast<volatile int*>(smem + 150 * 1024);
        int idx = 0;
        while (*(volatile int*)&stop == 0) {
          idx = p[idx & 63] & 63;
          ++cnt;
        }
        if (idx == 1234567) cnt = 0;
      }
      __syncwarp();
      if (tid == 32) out[1] = cnt;
    } else {
      // epilogue-like TMEM traffic: x16 loads over columns the MMAs do not touch
      const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384 + ((warp >> 2) & 3) * 16;
      uint32_t acc = 0;
      long long cnt = 0;
      while (*(volatile int*)&stop == 0) {
        uint32_t v[16];
        tmem_ld16(base, v);
        tmem_wait_ld();
        acc += v[0] ^ v[15];
        ++cnt;
      }
      if (tid == 32) out[1] = cnt + (acc == 0x12345u);
    }
  } else if (store_warps > 0 && warp <= store_warps) {
    // staging-like traffic: every lane stores 128-bit values to its own 16-byte slot pattern in a separate 32 KB region
    float4* dst = reinterpret_cast<float4*>(smem + 128 * 1024) + (warp - 1) * 128 + (tid & 31);
    float4 v = make_float4(1.f, 2.f, 3.f, (float)tid);
    long long cnt = 0;
    while (*(volatile int*)&stop == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j * 32] = v;
      v.x += 1.f;
      ++cnt;
    }
    if (tid == 32) out[1] = cnt;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  long long h[4];
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Case { int form, n, ks; const char* what; };
  const Case cases[] = {{0, 64, 8, "TS M=128 N=64 B MN-major (recompute), 24 MMAs"},
                        {0, 64, 2, "TS M=128 N=64 B MN-major K=16 (layer 0), 6 MMAs"},
                        {1, 64, 8, "TS M=128 N=64 B K-major (dgrad), 24 MMAs"},
                        {1, 16, 8, "TS M=128 N=16 B K-major (dgrad layer 0), 24 MMAs"},
                        {2, 64, 8, "SS M=64 N=64 rows=64 (wgrad half), 24 MMAs"},
                        {2, 64, 16, "SS M=64 N=64 rows=128 (wgrad full), 48 MMAs"},
                        {2, 16, 16, "SS M=64 N=16 rows=128 (wgrad layer 0), 48 MMAs"},
                        {3, 64, 16, "SS M=128 N=64 rows=128, 48 MMAs"},
                        {3, 128, 16, "SS M=128 N=128 rows=128, 48 MMAs"}};
  for (const Case& c : cases) {
    for (int sw : {0, -1501, -1503, -103}) {
      const int reps = 32;
      cudaMemset(d, 0, 64);
      rate_kernel<<<1, 512, 180 * 1024>>>(c.form, c.n, c.ks, reps, sw, d);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
      const int nmma = 3 * c.ks;
      printf("%-52s %-22s: %7.1f cycles/batch, %5.1f per MMA", c.what,
             sw == 0 ? "alone" : (sw > 0 ? "15 warps st.shared" : (sw == -1501 ? "15 warps poll mbarrier" : (sw == -1503 ? "15 warps ld.shared chain" : (sw == -103 ? "1 warp ld.shared chain" : "15 warps tcgen05.ld")))),
             (double)h[0] / reps, (double)h[0] / reps / nmma);
      if (sw > 0) printf("; concurrent stores: %6.1f B/cycle/SM", (double)h[1] * 4 * 32 * 16 * sw / (double)h[0]);
      if (sw == -1501) printf("; %6.1f cycles per poll", (double)h[0] / ((double)h[1] / 2 + 1e-9));
      if (sw == -1503 || sw == -103) printf("; %6.1f cycles per dependent ld.shared", (double)h[0] / ((double)h[1] + 1e-9));
      if (sw == -1502) printf("; concurrent ld: %6.1f B/cycle/SM", (double)h[1] * 32 * 64 * 15 / (double)h[0]);
      printf("  (%s)\n", cudaGetErrorString(e));
    }
  }
  return 0;
}
