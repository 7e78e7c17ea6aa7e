// Graph handle: COO -> stable dst-sorted CSR, edge permutation, src-sorted transpose, work units, and the
// merged (dst, ascending src) adjacency GCNConv's sparse-matmul semantics need.  Built once per graph on the
// device and cached by the caller behind `updategraph` -- it replaces the per-call index work of
// GraphNeuralNetworks.propagate and of GCNConv's add_self_loops/degree/adjacency_matrix
// (/root/reference/src/layers.jl:210-225, /root/reference/src/utils.jl:24-31).
#include <cub/cub.cuh>

#include <cstdarg>
#include <string>
#include <vector>

#include "ngpde_common.cuh"

namespace ngpde {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

namespace {

template <class T>
__global__ void convert_idx_kernel(const T* __restrict__ in, int* __restrict__ out, int64_t n, int64_t base, int64_t N,
                                   int* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t v = (int64_t)in[i] - base;
  if (v < 0 || v >= N) atomicExch(bad, 1);
  out[i] = (int)v;
}

__global__ void iota_kernel(int* out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int)i;
}

__global__ void gather_int_kernel(const int* __restrict__ in, const int* __restrict__ idx, int* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}

// ptr[j] = first position in the non-decreasing array `keys` (length n) with keys[pos] >= j, j = 0..N
__global__ void lower_bound_kernel(const int* __restrict__ keys, int64_t n, int* __restrict__ ptr, int64_t N) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j > N) return;
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (keys[mid] < j) lo = mid + 1; else hi = mid;
  }
  ptr[j] = (int)lo;
}

__global__ void gcn_keys_kernel(const int* __restrict__ s, const int* __restrict__ t, int64_t E, int64_t M, int64_t N,
                                unsigned long long* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  if (i < E) keys[i] = (unsigned long long)t[i] * (unsigned long long)N + (unsigned long long)s[i];
  else keys[i] = (unsigned long long)(i - E) * (unsigned long long)N + (unsigned long long)(i - E);
}

__global__ void head_flags_kernel(const unsigned long long* __restrict__ keys, int64_t M, int* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < M) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void gcn_fill_kernel(const unsigned long long* __restrict__ keys, const int* __restrict__ incl, int64_t M,
                                int64_t N, int nnz, int* __restrict__ runptr, int* __restrict__ rowval,
                                int* __restrict__ colidx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) runptr[nnz] = (int)M;
  if (i >= M) return;
  if (i == 0 || keys[i] != keys[i - 1]) {
    const int slot = incl[i] - 1;
    runptr[slot] = (int)i;
    rowval[slot] = (int)(keys[i] % (unsigned long long)N);
    colidx[slot] = (int)(keys[i] / (unsigned long long)N);
  }
}

int bits_for(unsigned long long maxval) {
  int b = 1;
  while (b < 64 && (maxval >> b) != 0) ++b;
  return b;
}

inline unsigned blocks_for(int64_t n, int t = 256) { return (unsigned)std::max<int64_t>(1, (n + t - 1) / t); }

struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  int alloc(size_t bytes) {
    NGPDE_CUDA_TRY(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
    return NGPDE_OK;
  }
};

template <class K>
int sort_pairs(const K* kin, K* kout, const int* vin, int* vout, int64_t n, int end_bit, cudaStream_t st) {
  size_t tb = 0;
  NGPDE_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tb, kin, kout, vin, vout, n, 0, end_bit, st));
  DevBuf tmp;
  if (int rc = tmp.alloc(tb)) return rc;
  NGPDE_CUDA_TRY(cub::DeviceRadixSort::SortPairs(tmp.p, tb, kin, kout, vin, vout, n, 0, end_bit, st));
  NGPDE_CUDA_TRY(cudaStreamSynchronize(st));
  return NGPDE_OK;
}

int alloc_int(int** p, int64_t n) {
  NGPDE_CUDA_TRY(cudaMalloc(p, sizeof(int) * std::max<int64_t>(n, 4)));
  return NGPDE_OK;
}

// Greedy packing of consecutive destination rows into units of at most `te` edges / `te` rows.
std::vector<int> greedy_units(const std::vector<int>& rowptr, int te) {
  const int n = (int)rowptr.size() - 1;
  std::vector<int> b;
  b.push_back(0);
  int i = 0;
  while (i < n) {
    int j = i + 1;
    while (j < n && j - i < te && rowptr[j + 1] - rowptr[i] <= te) ++j;
    b.push_back(j);
    i = j;
  }
  return b;
}

}  // namespace

int build_gcn_layout(ngpde_graph* g, int with_loops, cudaStream_t st) {
  GcnLayout& L = g->gcn[with_loops ? 1 : 0];
  if (L.built) return NGPDE_OK;
  const int64_t N = g->N, E = g->E;
  const int64_t M = E + (with_loops ? N : 0);
  if (int rc = alloc_int(&L.colptr, N + 1)) return rc;
  if (int rc = alloc_int(&L.tptr, N + 1)) return rc;
  if (int rc = alloc_int(&L.order, M)) return rc;
  if (M == 0) {
    NGPDE_CUDA_TRY(cudaMemsetAsync(L.colptr, 0, sizeof(int) * (N + 1), st));
    NGPDE_CUDA_TRY(cudaMemsetAsync(L.tptr, 0, sizeof(int) * (N + 1), st));
    if (int rc = alloc_int(&L.rowval, 1)) return rc;
    if (int rc = alloc_int(&L.colidx, 1)) return rc;
    if (int rc = alloc_int(&L.runptr, 1)) return rc;
    if (int rc = alloc_int(&L.tpos, 1)) return rc;
    NGPDE_CUDA_TRY(cudaMemsetAsync(L.runptr, 0, sizeof(int), st));
    L.nnz = 0;
    L.built = true;
    return NGPDE_OK;
  }
  DevBuf keys, keys_sorted, iota, flags, incl, rv_sorted, iota2;
  if (int rc = keys.alloc(sizeof(unsigned long long) * M)) return rc;
  if (int rc = keys_sorted.alloc(sizeof(unsigned long long) * M)) return rc;
  if (int rc = iota.alloc(sizeof(int) * M)) return rc;
  if (int rc = flags.alloc(sizeof(int) * M)) return rc;
  if (int rc = incl.alloc(sizeof(int) * M)) return rc;
  gcn_keys_kernel<<<blocks_for(M), 256, 0, st>>>(g->s_orig, g->t_orig, E, M, N, (unsigned long long*)keys.p);
  iota_kernel<<<blocks_for(M), 256, 0, st>>>((int*)iota.p, M);
  const int kb = bits_for((unsigned long long)N * (unsigned long long)N);
  if (int rc = sort_pairs((const unsigned long long*)keys.p, (unsigned long long*)keys_sorted.p, (const int*)iota.p,
                          L.order, M, kb, st))
    return rc;
  head_flags_kernel<<<blocks_for(M), 256, 0, st>>>((const unsigned long long*)keys_sorted.p, M, (int*)flags.p);
  {
    size_t tb = 0;
    NGPDE_CUDA_TRY(cub::DeviceScan::InclusiveSum(nullptr, tb, (const int*)flags.p, (int*)incl.p, M, st));
    DevBuf tmp;
    if (int rc = tmp.alloc(tb)) return rc;
    NGPDE_CUDA_TRY(cub::DeviceScan::InclusiveSum(tmp.p, tb, (const int*)flags.p, (int*)incl.p, M, st));
    NGPDE_CUDA_TRY(cudaStreamSynchronize(st));
  }
  int nnz = 0;
  NGPDE_CUDA_TRY(cudaMemcpy(&nnz, (const int*)incl.p + (M - 1), sizeof(int), cudaMemcpyDeviceToHost));
  L.nnz = nnz;
  if (int rc = alloc_int(&L.rowval, nnz)) return rc;
  if (int rc = alloc_int(&L.runptr, nnz + 1)) return rc;
  if (int rc = alloc_int(&L.tpos, nnz)) return rc;
  if (int rc = alloc_int(&L.colidx, nnz)) return rc;
  gcn_fill_kernel<<<blocks_for(M), 256, 0, st>>>((const unsigned long long*)keys_sorted.p, (const int*)incl.p, M, N,
                                                 nnz, L.runptr, L.rowval, L.colidx);
  lower_bound_kernel<<<blocks_for(N + 1), 256, 0, st>>>((const int*)L.colidx, nnz, L.colptr, N);
  // transpose: merged entries grouped by source, ascending destination (stable)
  if (int rc = rv_sorted.alloc(sizeof(int) * nnz)) return rc;
  if (int rc = iota2.alloc(sizeof(int) * nnz)) return rc;
  iota_kernel<<<blocks_for(nnz), 256, 0, st>>>((int*)iota2.p, nnz);
  if (int rc = sort_pairs((const int*)L.rowval, (int*)rv_sorted.p, (const int*)iota2.p, L.tpos, nnz,
                          bits_for((unsigned long long)std::max<int64_t>(N - 1, 1)), st))
    return rc;
  lower_bound_kernel<<<blocks_for(N + 1), 256, 0, st>>>((const int*)rv_sorted.p, nnz, L.tptr, N);
  NGPDE_CUDA_TRY(cudaStreamSynchronize(st));
  NGPDE_CUDA_TRY(cudaGetLastError());
  L.built = true;
  return NGPDE_OK;
}

}  // namespace ngpde

using namespace ngpde;

extern "C" int ngpde_version(void) { return NGPDE_VERSION; }
extern "C" const char* ngpde_last_error(void) { return g_last_error.c_str(); }

extern "C" int ngpde_graph_create(ngpde_graph_t* out, int64_t num_nodes, int64_t num_edges, const void* src,
                                  const void* dst, int32_t index_dtype, int32_t index_base, int32_t indices_on_device,
                                  int64_t num_graphs, void* stream) {
  NGPDE_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  NGPDE_REQUIRE(num_nodes >= 0 && num_edges >= 0, "negative graph size");
  NGPDE_REQUIRE(num_nodes < (int64_t(1) << 31) - 1 && num_edges < (int64_t(1) << 31) - 1, "graph too large for int32 indices");
  NGPDE_REQUIRE(num_edges == 0 || (src != nullptr && dst != nullptr), "src/dst are NULL");
  NGPDE_REQUIRE(index_dtype == NGPDE_IDX_I32 || index_dtype == NGPDE_IDX_I64, "unknown index dtype %d", index_dtype);
  NGPDE_REQUIRE(index_base == 0 || index_base == 1, "index_base must be 0 or 1");
  NGPDE_REQUIRE(num_graphs >= 1, "num_graphs must be >= 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t N = num_nodes, E = num_edges;

  ngpde_graph* g = new ngpde_graph();
  g->N = N; g->E = E; g->G = num_graphs;
  struct Guard {
    ngpde_graph* g;
    ~Guard() { if (g) ngpde_graph_destroy(g); }
  } guard{g};
  NGPDE_CUDA_TRY(cudaGetDevice(&g->device));
  NGPDE_CUDA_TRY(cudaDeviceGetAttribute(&g->num_sms, cudaDevAttrMultiProcessorCount, g->device));

  if (int rc = alloc_int(&g->s_orig, E)) return rc;
  if (int rc = alloc_int(&g->t_orig, E)) return rc;
  if (int rc = alloc_int(&g->rowptr, N + 1)) return rc;
  if (int rc = alloc_int(&g->src, E)) return rc;
  if (int rc = alloc_int(&g->dst, E)) return rc;
  if (int rc = alloc_int(&g->perm, E)) return rc;
  if (int rc = alloc_int(&g->tptr, N + 1)) return rc;
  if (int rc = alloc_int(&g->tpos, E)) return rc;

  std::vector<int> h_rowptr(N + 1, 0);
  if (E > 0) {
    const size_t isz = index_dtype == NGPDE_IDX_I64 ? 8 : 4;
    DevBuf stage_s, stage_t, bad, iota, skeys;
    const void* ds = src;
    const void* dt = dst;
    if (!indices_on_device) {
      if (int rc = stage_s.alloc(isz * E)) return rc;
      if (int rc = stage_t.alloc(isz * E)) return rc;
      NGPDE_CUDA_TRY(cudaMemcpyAsync(stage_s.p, src, isz * E, cudaMemcpyHostToDevice, st));
      NGPDE_CUDA_TRY(cudaMemcpyAsync(stage_t.p, dst, isz * E, cudaMemcpyHostToDevice, st));
      ds = stage_s.p;
      dt = stage_t.p;
    }
    if (int rc = bad.alloc(sizeof(int))) return rc;
    NGPDE_CUDA_TRY(cudaMemsetAsync(bad.p, 0, sizeof(int), st));
    if (index_dtype == NGPDE_IDX_I64) {
      convert_idx_kernel<long long><<<blocks_for(E), 256, 0, st>>>((const long long*)ds, g->s_orig, E, index_base, N, (int*)bad.p);
      convert_idx_kernel<long long><<<blocks_for(E), 256, 0, st>>>((const long long*)dt, g->t_orig, E, index_base, N, (int*)bad.p);
    } else {
      convert_idx_kernel<int><<<blocks_for(E), 256, 0, st>>>((const int*)ds, g->s_orig, E, index_base, N, (int*)bad.p);
      convert_idx_kernel<int><<<blocks_for(E), 256, 0, st>>>((const int*)dt, g->t_orig, E, index_base, N, (int*)bad.p);
    }
    int h_bad = 0;
    NGPDE_CUDA_TRY(cudaMemcpyAsync(&h_bad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    NGPDE_CUDA_TRY(cudaStreamSynchronize(st));
    NGPDE_REQUIRE(h_bad == 0, "edge index out of range [%d, %lld]", (int)index_base, (long long)(N - 1 + index_base));

    const int nb = bits_for((unsigned long long)std::max<int64_t>(N - 1, 1));
    if (int rc = iota.alloc(sizeof(int) * E)) return rc;
    iota_kernel<<<blocks_for(E), 256, 0, st>>>((int*)iota.p, E);
    // stable sort by destination: within a row, ascending original position == NNlib.scatter's visiting order
    if (int rc = sort_pairs((const int*)g->t_orig, g->dst, (const int*)iota.p, g->perm, E, nb, st)) return rc;
    gather_int_kernel<<<blocks_for(E), 256, 0, st>>>(g->s_orig, g->perm, g->src, E);
    lower_bound_kernel<<<blocks_for(N + 1), 256, 0, st>>>(g->dst, E, g->rowptr, N);
    // transpose: CSR positions grouped by source (stable)
    if (int rc = skeys.alloc(sizeof(int) * E)) return rc;
    if (int rc = sort_pairs((const int*)g->src, (int*)skeys.p, (const int*)iota.p, g->tpos, E, nb, st)) return rc;
    lower_bound_kernel<<<blocks_for(N + 1), 256, 0, st>>>((const int*)skeys.p, E, g->tptr, N);
    NGPDE_CUDA_TRY(cudaMemcpyAsync(h_rowptr.data(), g->rowptr, sizeof(int) * (N + 1), cudaMemcpyDeviceToHost, st));
    NGPDE_CUDA_TRY(cudaStreamSynchronize(st));
  } else {
    NGPDE_CUDA_TRY(cudaMemsetAsync(g->rowptr, 0, sizeof(int) * (N + 1), st));
    NGPDE_CUDA_TRY(cudaMemsetAsync(g->tptr, 0, sizeof(int) * (N + 1), st));
  }
  for (int t = 0; t < 3; ++t) {
    std::vector<int> u = greedy_units(h_rowptr, kTileSizes[t]);
    g->n_units[t] = (int)u.size() - 1;
    if (int rc = alloc_int(&g->units[t], (int64_t)u.size())) return rc;
    NGPDE_CUDA_TRY(cudaMemcpyAsync(g->units[t], u.data(), sizeof(int) * u.size(), cudaMemcpyHostToDevice, st));
    NGPDE_CUDA_TRY(cudaStreamSynchronize(st));
  }
  NGPDE_CUDA_TRY(cudaGetLastError());
  guard.g = nullptr;
  *out = g;
  return NGPDE_OK;
}

extern "C" int ngpde_graph_destroy(ngpde_graph_t g) {
  if (!g) return NGPDE_OK;
  int* ptrs[] = {g->s_orig, g->t_orig, g->rowptr, g->src, g->dst, g->perm, g->tptr, g->tpos,
                 g->units[0], g->units[1], g->units[2]};
  for (int* p : ptrs) if (p) cudaFree(p);
  for (auto& L : g->gcn) {
    int* q[] = {L.colptr, L.rowval, L.colidx, L.runptr, L.order, L.tptr, L.tpos};
    for (int* p : q) if (p) cudaFree(p);
  }
  delete g;
  return NGPDE_OK;
}

extern "C" int64_t ngpde_graph_num_nodes(ngpde_graph_t g) { return g ? g->N : -1; }
extern "C" int64_t ngpde_graph_num_edges(ngpde_graph_t g) { return g ? g->E : -1; }

extern "C" int ngpde_graph_array(ngpde_graph_t g, int32_t which, int32_t with_self_loops, const void** device_ptr,
                                 int64_t* len) {
  NGPDE_REQUIRE(g && device_ptr && len, "null argument");
  const int* p = nullptr;
  int64_t n = 0;
  if (which >= NGPDE_GA_GCN_COLPTR) {
    if (int rc = build_gcn_layout(g, with_self_loops, 0)) return rc;
  }
  const GcnLayout& L = g->gcn[with_self_loops ? 1 : 0];
  switch (which) {
    case NGPDE_GA_ROWPTR: p = g->rowptr; n = g->N + 1; break;
    case NGPDE_GA_SRC: p = g->src; n = g->E; break;
    case NGPDE_GA_DST: p = g->dst; n = g->E; break;
    case NGPDE_GA_PERM: p = g->perm; n = g->E; break;
    case NGPDE_GA_TPTR: p = g->tptr; n = g->N + 1; break;
    case NGPDE_GA_TPOS: p = g->tpos; n = g->E; break;
    case NGPDE_GA_UNITS32: p = g->units[0]; n = g->n_units[0] + 1; break;
    case NGPDE_GA_UNITS64: p = g->units[1]; n = g->n_units[1] + 1; break;
    case NGPDE_GA_UNITS128: p = g->units[2]; n = g->n_units[2] + 1; break;
    case NGPDE_GA_GCN_COLPTR: p = L.colptr; n = g->N + 1; break;
    case NGPDE_GA_GCN_ROWVAL: p = L.rowval; n = L.nnz; break;
    case NGPDE_GA_GCN_TPTR: p = L.tptr; n = g->N + 1; break;
    case NGPDE_GA_GCN_TPOS: p = L.tpos; n = L.nnz; break;
    default: set_error("unknown graph array %d", which); return NGPDE_ERR_INVALID;
  }
  *device_ptr = p;
  *len = n;
  return NGPDE_OK;
}

extern "C" int ngpde_graph_array_copy(ngpde_graph_t g, int32_t which, int32_t with_self_loops, int32_t* dst_device,
                                      int64_t capacity, void* stream) {
  const void* p = nullptr;
  int64_t n = 0;
  if (int rc = ngpde_graph_array(g, which, with_self_loops, &p, &n)) return rc;
  NGPDE_REQUIRE(capacity >= n && (n == 0 || dst_device != nullptr), "destination too small: %lld < %lld",
                (long long)capacity, (long long)n);
  if (n > 0)
    NGPDE_CUDA_TRY(cudaMemcpyAsync(dst_device, p, sizeof(int) * n, cudaMemcpyDeviceToDevice,
                                   static_cast<cudaStream_t>(stream)));
  return NGPDE_OK;
}
