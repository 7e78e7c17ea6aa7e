// Cost of one thread-block-cluster barrier (8 CTAs x 512 threads), with and without a DSMEM store before it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/cluster_sync_bench tools/cluster_sync_bench.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(512, 1) k(int reps, int mode, long long* out, float* g) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float buf[1024];
  const int tid = threadIdx.x;
  cluster.sync();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (mode == 1 && tid < 128) {
      for (int c = 0; c < 8; ++c) cluster.map_shared_rank(buf, c)[blockIdx.x * 128 + tid] = (float)r;
    }
    if (mode == 2) {
      g[blockIdx.x * 512 + tid] = (float)r;
      __threadfence();
    }
    if (mode == 3) __syncthreads();
    else cluster.sync();
  }
  const long long t1 = clock64();
  if (tid == 0 && blockIdx.x == 0) out[0] = (t1 - t0) / reps;
  if (buf[tid] == 123.f) out[1] = 1;
}

int main() {
  long long* d; float* g;
  cudaMalloc(&d, 64); cudaMalloc(&g, 1 << 16);
  const char* names[] = {"cluster.sync only", "DSMEM broadcast (128 floats x 8) + cluster.sync", "global store + __threadfence + cluster.sync", "__syncthreads only"};
  for (int mode = 0; mode < 4; ++mode) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(8); cfg.blockDim = dim3(512);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int reps = 1000;
    cudaLaunchKernelEx(&cfg, k, reps, mode, d, g);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-52s %6lld cycles per iteration (%s)\n", names[mode], h[0], cudaGetErrorString(e));
  }
  return 0;
}
