// The training step either side of the adjoint: loss + cotangent kernels and the optimiser update
// over the flat parameter vector, so that one training iteration never leaves the device.
//
// The reference calls these through un-vendored packages (Project.toml: Optimisers, Flux.Losses via the tutorials):
//   * Optimisers.Adam(0.01f0)                      docs/src/tutorials/graph_node.md:122-129
//   * Optimisers.Rprop(1f-6, (0.5f0, 1.2f0), (1f-8, 10f0))   docs/src/tutorials/VMH.md:97
//   * mse(y^, y) = mean(abs2, y^ - y)              docs/src/tutorials/VMH.md:105-109
//   * logitcrossentropy(y^[:, mask], y) = mean(-sum(y .* logsoftmax(y^); dims = 1))   graph_node.md:100-106
// Their published update rules are restated in oracle/ngpde_oracle.py; the element-wise kernels below use the same
// operation order with individually rounded IEEE operations (no FMA contraction), so the parameter updates are bit-exact
// against the float32 restatement.  Reductions are two-stage in a fixed order: deterministic, no atomics.
#include <algorithm>
#include <vector>

#include "ngpde_common.cuh"

namespace ngpde {
namespace {

// Optimisers.jl Adam:  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  x -= m / (1 - b1^t) / (sqrt(v / (1 - b2^t)) + eps) * eta
__global__ void adam_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float eta, float b1, float b2, float eps, float b1t, float b2t) {
  const float omb1 = __fsub_rn(1.f, b1), omb2 = __fsub_rn(1.f, b2);
  const float c1 = __fsub_rn(1.f, b1t), c2 = __fsub_rn(1.f, b2t);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = __fadd_rn(__fmul_rn(b1, m[i]), __fmul_rn(omb1, gi));
    const float vi = __fadd_rn(__fmul_rn(b2, v[i]), __fmul_rn(omb2, __fmul_rn(gi, gi)));
    m[i] = mi;
    v[i] = vi;
    const float den = __fadd_rn(__fsqrt_rn(__fdiv_rn(vi, c2)), eps);
    const float step = __fmul_rn(__fdiv_rn(__fdiv_rn(mi, c1), den), eta);
    x[i] = __fsub_rn(x[i], step);
  }
}

// Optimisers.jl Rprop: per-element step size eta_i grows by l+ while the gradient keeps its sign, shrinks by l- (and the
// stored gradient is zeroed) when it flips;  x -= eta_i * sign(g_i)
__global__ void rprop_kernel(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ gprev,
                             float* __restrict__ eta, long long n, float lminus, float lplus, float gmin, float gmax) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float dx = g[i], gp = gprev[i];
    const float prod = __fmul_rn(gp, dx);
    float e = eta[i];
    if (prod > 0.f) e = fminf(__fmul_rn(e, lplus), gmax);
    else if (prod < 0.f) e = fmaxf(__fmul_rn(e, lminus), gmin);
    const float gn = prod < 0.f ? 0.f : dx;
    eta[i] = e;
    gprev[i] = gn;
    const float sgn = gn > 0.f ? 1.f : (gn < 0.f ? -1.f : 0.f);
    x[i] = __fsub_rn(x[i], __fmul_rn(e, sgn));
  }
}

constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 1024;

// fixed-order block reduction: thread-strided partial sums, then a shared-memory tree
__device__ __forceinline__ float block_sum(float v, float* sm) {
  sm[threadIdx.x] = v;
  __syncthreads();
  for (int s = RED_THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sm[threadIdx.x] = sm[threadIdx.x] + sm[threadIdx.x + s];
    __syncthreads();
  }
  return sm[0];
}

// stage 1 of mean(abs2, yhat - y): partial[b] = sum over the block's grid-stride elements; also the cotangent 2 (yhat - y) / n
__global__ void __launch_bounds__(RED_THREADS) mse_stage1_kernel(const float* __restrict__ yhat, const float* __restrict__ y,
                                                                 long long n, float scale, float* __restrict__ dyhat,
                                                                 float* __restrict__ partial) {
  __shared__ float sm[RED_THREADS];
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = yhat[i] - y[i];
    acc += d * d;
    if (dyhat) dyhat[i] = scale * d;
  }
  const float s = block_sum(acc, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// stage 2: loss = (sum of partials in ascending block order) * inv_n
__global__ void reduce_scalar_kernel(const float* __restrict__ partial, int nblk, float inv_n, float* __restrict__ out) {
  __shared__ float sm[RED_THREADS];
  float acc = 0.f;
  for (int i = threadIdx.x; i < nblk; i += RED_THREADS) acc += partial[i];
  const float s = block_sum(acc, sm);
  if (threadIdx.x == 0) *out = s * inv_n;
}

// one thread per masked column j: logsoftmax over the C classes of row idx[j] of yhat, loss_j = -sum_c y[j][c] lsm_c,
// cotangent dyhat[row][c] = (softmax_c * sum_c' y[j][c'] - y[j][c]) / nm.  dyhat must have been zeroed (rows outside the mask).
__global__ void __launch_bounds__(RED_THREADS) ce_stage1_kernel(const float* __restrict__ yhat, const float* __restrict__ y,
                                                                const int* __restrict__ idx, long long nm, int C,
                                                                float inv_nm, float* __restrict__ dyhat,
                                                                float* __restrict__ partial) {
  __shared__ float sm[RED_THREADS];
  float acc = 0.f;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nm; j += (long long)gridDim.x * blockDim.x) {
    const long long row = idx ? (long long)idx[j] : j;
    const float* z = yhat + row * C;
    const float* t = y + j * C;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, z[c]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(z[c] - mx);
    const float lse = logf(se);
    float l = 0.f, ty = 0.f;
    for (int c = 0; c < C; ++c) {
      l -= t[c] * ((z[c] - mx) - lse);
      ty += t[c];
    }
    acc += l;
    if (dyhat) {
      float* dz = dyhat + row * C;
      for (int c = 0; c < C; ++c) dz[c] = (expf((z[c] - mx) - lse) * ty - t[c]) * inv_nm;
    }
  }
  const float s = block_sum(acc, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

int red_blocks(long long n) { return (int)std::max<long long>(1, std::min<long long>((n + RED_THREADS - 1) / RED_THREADS, RED_MAX_BLOCKS)); }

}  // namespace
}  // namespace ngpde

using namespace ngpde;

extern "C" int ngpde_adam_step(float* params, const float* grad, float* m, float* v, int64_t n, float eta, float beta1,
                               float beta2, float eps, float beta1_t, float beta2_t, void* stream) {
  NGPDE_REQUIRE(n >= 0 && (n == 0 || (params && grad && m && v)), "adam: null argument");
  if (n == 0) return NGPDE_OK;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  adam_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grad, m, v, n, eta, beta1, beta2, eps, beta1_t, beta2_t);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

extern "C" int ngpde_rprop_step(float* params, const float* grad, float* g_prev, float* eta, int64_t n, float ell_minus,
                                float ell_plus, float gamma_min, float gamma_max, void* stream) {
  NGPDE_REQUIRE(n >= 0 && (n == 0 || (params && grad && g_prev && eta)), "rprop: null argument");
  if (n == 0) return NGPDE_OK;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  rprop_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(params, grad, g_prev, eta, n, ell_minus, ell_plus, gamma_min, gamma_max);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

extern "C" size_t ngpde_loss_workspace_bytes(void) { return sizeof(float) * RED_MAX_BLOCKS; }

extern "C" int ngpde_mse_loss(const float* yhat, const float* y, int64_t n, float* loss, float* dyhat, void* workspace,
                              size_t workspace_bytes, void* stream) {
  NGPDE_REQUIRE(yhat && y && loss && n > 0, "mse: bad argument");
  NGPDE_REQUIRE(workspace && workspace_bytes >= ngpde_loss_workspace_bytes(), "mse: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  const int nb = red_blocks(n);
  mse_stage1_kernel<<<nb, RED_THREADS, 0, st>>>(yhat, y, n, 2.f / (float)n, dyhat, partial);
  reduce_scalar_kernel<<<1, RED_THREADS, 0, st>>>(partial, nb, 1.f / (float)n, loss);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

extern "C" int ngpde_logit_cross_entropy(const float* yhat, int64_t n_rows, int32_t n_classes, const float* y,
                                         const int32_t* mask_idx, int64_t n_masked, float* loss, float* dyhat,
                                         void* workspace, size_t workspace_bytes, void* stream) {
  NGPDE_REQUIRE(yhat && y && loss && n_classes > 0 && n_masked > 0 && n_rows >= n_masked, "logitcrossentropy: bad argument");
  NGPDE_REQUIRE(mask_idx || n_masked == n_rows, "logitcrossentropy: mask_idx is NULL but n_masked != n_rows");
  NGPDE_REQUIRE(workspace && workspace_bytes >= ngpde_loss_workspace_bytes(), "logitcrossentropy: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  if (dyhat) NGPDE_CUDA_TRY(cudaMemsetAsync(dyhat, 0, sizeof(float) * (size_t)n_rows * n_classes, st));
  const int nb = red_blocks(n_masked);
  ce_stage1_kernel<<<nb, RED_THREADS, 0, st>>>(yhat, y, mask_idx, n_masked, n_classes, 1.f / (float)n_masked, dyhat, partial);
  reduce_scalar_kernel<<<1, RED_THREADS, 0, st>>>(partial, nb, 1.f / (float)n_masked, loss);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

// Number of kernel nodes (and of all nodes) of a captured CUDA graph: how bench.py counts, exactly, the kernels one
// replayed step launches (`gpu_launches`) instead of keeping a hand-maintained table.
extern "C" int ngpde_cuda_graph_kernel_nodes(void* cuda_graph, int64_t* n_kernels, int64_t* n_nodes) {
  NGPDE_REQUIRE(cuda_graph && n_kernels && n_nodes, "null argument");
  cudaGraph_t g = static_cast<cudaGraph_t>(cuda_graph);
  size_t n = 0;
  NGPDE_CUDA_TRY(cudaGraphGetNodes(g, nullptr, &n));
  std::vector<cudaGraphNode_t> nodes(n);
  if (n) NGPDE_CUDA_TRY(cudaGraphGetNodes(g, nodes.data(), &n));
  int64_t k = 0;
  for (size_t i = 0; i < n; ++i) {
    cudaGraphNodeType t;
    NGPDE_CUDA_TRY(cudaGraphNodeGetType(nodes[i], &t));
    if (t == cudaGraphNodeTypeKernel) ++k;
  }
  *n_kernels = k;
  *n_nodes = (int64_t)n;
  return NGPDE_OK;
}
