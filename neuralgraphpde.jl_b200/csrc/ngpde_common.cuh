// Internal declarations shared by the libngpde translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "ngpde.h"

namespace ngpde {

void set_error(const char* fmt, ...);

#define NGPDE_CUDA_TRY(expr)                                                                      \
  do {                                                                                            \
    cudaError_t err__ = (expr);                                                                   \
    if (err__ != cudaSuccess) {                                                                   \
      ::ngpde::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(err__)); \
      return NGPDE_ERR_CUDA;                                                                      \
    }                                                                                             \
  } while (0)

#define NGPDE_REQUIRE(cond, ...)          \
  do {                                    \
    if (!(cond)) {                        \
      ::ngpde::set_error(__VA_ARGS__);    \
      return NGPDE_ERR_INVALID;           \
    }                                     \
  } while (0)

constexpr int kTileSizes[3] = {32, 64, 128};

// Merged (dst, ascending src) adjacency used by GCNConv, with or without appended self-loops.
struct GcnLayout {
  bool built = false;
  int nnz = 0;
  int* colptr = nullptr;  // [N+1]
  int* rowval = nullptr;  // [nnz] source of each merged entry
  int* colidx = nullptr;  // [nnz] destination of each merged entry
  int* runptr = nullptr;  // [nnz+1] range of `order` that merged into each entry
  int* order = nullptr;   // [E(+N)] original edge ids sorted by (dst, src), stable; ids >= E are self-loops
  int* tptr = nullptr;    // [N+1] transpose: entries grouped by source, ascending dst
  int* tpos = nullptr;    // [nnz]
};

}  // namespace ngpde

struct ngpde_graph {
  int64_t N = 0, E = 0, G = 1;
  int device = 0;
  int num_sms = 148;
  int* s_orig = nullptr;  // [E] 0-based int32 copies of the COO lists (original order)
  int* t_orig = nullptr;
  int* rowptr = nullptr;  // [N+1]
  int* src = nullptr;     // [E] CSR order
  int* dst = nullptr;     // [E]
  int* perm = nullptr;    // [E]
  int* tptr = nullptr;    // [N+1]
  int* tpos = nullptr;    // [E]
  int* units[3] = {nullptr, nullptr, nullptr};
  int n_units[3] = {0, 0, 0};
  ngpde::GcnLayout gcn[2];
  int ode_max_tedges[2] = {-1, -1};
  int ode_wide_cluster[2] = {-1, -1};  // [forward, adjoint]: can a 16-CTA cluster be scheduled on this device (-1: not probed yet)
  int ode_max_edges[2] = {-1, -1};  // largest in-edge count of the persistent ODE kernels' node ranges (computed on first use)
};

namespace ngpde {

int build_gcn_layout(ngpde_graph* g, int with_loops, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// activations (accurate float32: tanhf/expf, no fast-math -- the per-layer tolerance is 1e-5)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(int a, float x) {
  switch (a) {
    case NGPDE_ACT_RELU: return fmaxf(x, 0.f);
    case NGPDE_ACT_TANH: return tanhf(x);
    case NGPDE_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case NGPDE_ACT_SWISH: return x / (1.f + expf(-x));
    case NGPDE_ACT_GELU: {
      const float k = 0.7978845608028654f;
      return 0.5f * x * (1.f + tanhf(k * (x + 0.044715f * x * x * x)));
    }
    case NGPDE_ACT_SOFTPLUS: return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
    case NGPDE_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case NGPDE_ACT_LEAKYRELU: return x > 0.f ? x : 0.01f * x;
    default: return x;
  }
}

// true when act'(p) can be written as a function of y = act(p)
__host__ __device__ __forceinline__ bool act_grad_from_y(int a) {
  return a != NGPDE_ACT_SWISH && a != NGPDE_ACT_GELU;
}

__device__ __forceinline__ float act_grad_y(int a, float y) {
  switch (a) {
    case NGPDE_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case NGPDE_ACT_TANH: return 1.f - y * y;
    case NGPDE_ACT_SIGMOID: return y * (1.f - y);
    case NGPDE_ACT_SOFTPLUS: return 1.f - expf(-y);
    case NGPDE_ACT_ELU: return y > 0.f ? 1.f : y + 1.f;
    case NGPDE_ACT_LEAKYRELU: return y > 0.f ? 1.f : 0.01f;
    default: return 1.f;
  }
}

__device__ __forceinline__ float act_grad_pre(int a, float p) {
  switch (a) {
    case NGPDE_ACT_SWISH: {
      float s = 1.f / (1.f + expf(-p));
      return s * (1.f + p * (1.f - s));
    }
    case NGPDE_ACT_GELU: {
      const float k = 0.7978845608028654f;
      float u = k * (p + 0.044715f * p * p * p);
      float t = tanhf(u);
      float du = k * (1.f + 3.f * 0.044715f * p * p);
      return 0.5f * (1.f + t) + 0.5f * p * (1.f - t * t) * du;
    }
    default: return act_grad_y(a, act_fwd(a, p));
  }
}

}  // namespace ngpde
