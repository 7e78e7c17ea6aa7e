// Latency / issue rate of the warp-level (legacy) tensor-core MMAs on sm_100a, as the persistent ODE kernels use them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/hmma_rate tools/hmma_rate.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

template <int KIND, int CHAINS>
__global__ void k(int reps, long long* out, float* sink) {
  float c[CHAINS][4];
  for (int j = 0; j < CHAINS; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
  uint32_t a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x * 5u, 11u};
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) {
      if (KIND == 0)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else if (KIND == 1)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
      else if (KIND == 2)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(b[0]));
      else
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3]) : "r"(a[0]), "r"(a[1]), "r"(b[0]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < CHAINS; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  if (s == 123.456f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int KIND, int CHAINS>
void run(const char* name, int threads, long long* d, float* sink) {
  const int reps = 2000;
  k<KIND, CHAINS><<<1, threads>>>(reps, d, sink);
  k<KIND, CHAINS><<<1, threads>>>(reps, d, sink);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / reps / CHAINS;
  printf("%-28s %4d threads (%d warps/SMSP) %d independent chains: %7.1f cycles per MMA per warp -> %6.1f cycles per MMA per SMSP\n", name, threads,
         threads / 128 > 0 ? threads / 128 : 1, CHAINS, per, per / (threads >= 128 ? threads / 128 : 1));
}

int main() {
  long long* d; float* sink;
  cudaMalloc(&d, 64); cudaMalloc(&sink, 64);
  run<0, 1>("m16n8k8 tf32", 32, d, sink);
  run<0, 4>("m16n8k8 tf32", 32, d, sink);
  run<0, 4>("m16n8k8 tf32", 128, d, sink);
  run<0, 4>("m16n8k8 tf32", 512, d, sink);
  run<2, 1>("m16n8k4 tf32", 32, d, sink);
  run<2, 4>("m16n8k4 tf32", 512, d, sink);
  run<1, 1>("m16n8k16 bf16", 32, d, sink);
  run<1, 4>("m16n8k16 bf16", 32, d, sink);
  run<1, 4>("m16n8k16 bf16", 128, d, sink);
  run<1, 4>("m16n8k16 bf16", 512, d, sink);
  run<3, 1>("m16n8k8 bf16", 32, d, sink);
  run<3, 4>("m16n8k8 bf16", 512, d, sink);
  return 0;
}
