// Halo-exchange helpers of the node-partitioned (single large graph) multi-GPU path -- SURVEY.md section 8e.
// The reference is single-device; these kernels are the device side of what a partitioned `propagate` needs around the
// NCCL exchange: packing the boundary rows a peer asked for, and adding the returned halo cotangents back into the
// owner's rows in a fixed order (no atomics, so the backward stays run-to-run deterministic).
//
//   ngpde_rows_gather       out[i][:] = x[rows[i]][:]                       (pack; also used for static halo data)
//   ngpde_rows_put          the same rows written straight into per-peer destination buffers (peer-mapped memory over
//                           NVLink: pack and transfer in one kernel, no intermediate send buffer)
//   ngpde_rows_segment_add  dst[seg_rows[u]][:] += sum_q src[seg_pos[q]][:], q ascending in [seg_ptr[u], seg_ptr[u+1])
#include <algorithm>

#include "ngpde_common.cuh"

namespace ngpde {
namespace {

// one thread per 16-byte chunk when d % 4 == 0, else per float
template <int V>
__global__ void rows_gather_kernel(const float* __restrict__ x, const int* __restrict__ rows, int64_t n_rows, int d,
                                   float* __restrict__ out) {
  const int dv = d / V;
  const int64_t total = n_rows * dv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dv;
    const int c = (int)(i - r * dv);
    const int64_t srow = rows[r];
    if (V == 4) {
      reinterpret_cast<float4*>(out)[r * dv + c] = reinterpret_cast<const float4*>(x)[srow * dv + c];
    } else {
      out[r * dv + c] = x[srow * dv + c];
    }
  }
}

template <int V>
__global__ void rows_put_kernel(const float* __restrict__ x, const int* __restrict__ rows,
                                const int64_t* __restrict__ peer_ptr, float* const* __restrict__ peer_dst, int n_peers,
                                int d) {
  const int dv = d / V;
  const int64_t total = peer_ptr[n_peers] * dv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dv;
    const int c = (int)(i - r * dv);
    int p = 0;
    while (p + 1 < n_peers && r >= peer_ptr[p + 1]) ++p;  // n_peers <= 8 on one node
    float* dstp = peer_dst[p];
    const int64_t lr = r - peer_ptr[p];
    const int64_t srow = rows[r];
    if (V == 4) {
      reinterpret_cast<float4*>(dstp)[lr * dv + c] = reinterpret_cast<const float4*>(x)[srow * dv + c];
    } else {
      dstp[lr * dv + c] = x[srow * dv + c];
    }
  }
}

__global__ void rows_segment_add_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                        const int* __restrict__ seg_rows, const int* __restrict__ seg_ptr,
                                        const int* __restrict__ seg_pos, int64_t n_segs, int d) {
  const int64_t total = n_segs * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t u = i / d;
    const int c = (int)(i - u * d);
    float acc = dst[(int64_t)seg_rows[u] * d + c];
    for (int q = seg_ptr[u]; q < seg_ptr[u + 1]; ++q) acc += src[(int64_t)seg_pos[q] * d + c];
    dst[(int64_t)seg_rows[u] * d + c] = acc;
  }
}

// One-shot all-reduce over peer-mapped buffers: every rank reads all `world` buffers (its own and its peers', over
// NVLink / NVSwitch) and adds them in ascending rank order -- the same order on every rank, so the result is deterministic
// and bit-identical across ranks.  Loads bypass L1 (the peers rewrite their buffers every step).
__global__ void peer_allreduce_kernel(const float* const* __restrict__ peers, int world, float* __restrict__ out, int64_t n) {
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 acc = __ldcv(reinterpret_cast<const float4*>(peers[0]) + i);
    for (int r = 1; r < world; ++r) {
      const float4 v = __ldcv(reinterpret_cast<const float4*>(peers[r]) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    float acc = __ldcv(peers[0] + i);
    for (int r = 1; r < world; ++r) acc += __ldcv(peers[r] + i);
    out[i] = acc;
  }
}

int grid_for(int64_t total) { return (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, 148 * 8)); }

}  // namespace
}  // namespace ngpde

using namespace ngpde;

extern "C" int ngpde_rows_gather(const float* x, const int32_t* rows, int64_t n_rows, int32_t d, float* out,
                                 void* stream) {
  NGPDE_REQUIRE(n_rows >= 0 && d > 0, "rows_gather: n_rows=%lld d=%d", (long long)n_rows, d);
  if (n_rows == 0) return NGPDE_OK;
  NGPDE_REQUIRE(x && rows && out, "rows_gather: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = d % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec) rows_gather_kernel<4><<<grid_for(n_rows * (d / 4)), 256, 0, st>>>(x, rows, n_rows, d, out);
  else rows_gather_kernel<1><<<grid_for(n_rows * d), 256, 0, st>>>(x, rows, n_rows, d, out);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

extern "C" int ngpde_rows_put(const float* x, const int32_t* rows, const int64_t* peer_ptr, float* const* peer_dst,
                              int32_t n_peers, int64_t n_rows, int32_t d, void* stream) {
  NGPDE_REQUIRE(n_peers >= 1 && n_peers <= 64 && d > 0 && n_rows >= 0, "rows_put: n_peers=%d d=%d", n_peers, d);
  if (n_rows == 0) return NGPDE_OK;
  NGPDE_REQUIRE(x && rows && peer_ptr && peer_dst, "rows_put: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // destination alignment is the caller's contract (buffers are 16-byte aligned and d % 4 == 0 keeps row starts aligned)
  const bool vec = d % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (vec) rows_put_kernel<4><<<grid_for(n_rows * (d / 4)), 256, 0, st>>>(x, rows, peer_ptr, peer_dst, n_peers, d);
  else rows_put_kernel<1><<<grid_for(n_rows * d), 256, 0, st>>>(x, rows, peer_ptr, peer_dst, n_peers, d);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

extern "C" int ngpde_rows_segment_add(float* dst, const float* src, const int32_t* seg_rows, const int32_t* seg_ptr,
                                      const int32_t* seg_pos, int64_t n_segs, int32_t d, void* stream) {
  NGPDE_REQUIRE(n_segs >= 0 && d > 0, "rows_segment_add: n_segs=%lld d=%d", (long long)n_segs, d);
  if (n_segs == 0) return NGPDE_OK;
  NGPDE_REQUIRE(dst && src && seg_rows && seg_ptr && seg_pos, "rows_segment_add: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rows_segment_add_kernel<<<grid_for(n_segs * d), 256, 0, st>>>(dst, src, seg_rows, seg_ptr, seg_pos, n_segs, d);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

extern "C" int ngpde_peer_allreduce_sum(const float* const* peer_bufs, int32_t world, float* out, int64_t n, void* stream) {
  NGPDE_REQUIRE(peer_bufs && out && world >= 1 && n >= 0, "peer_allreduce_sum: bad argument");
  NGPDE_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "peer_allreduce_sum: out must be 16-byte aligned");
  if (n == 0) return NGPDE_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  peer_allreduce_kernel<<<grid_for((n + 3) / 4), 256, 0, st>>>(peer_bufs, world, out, n);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}
