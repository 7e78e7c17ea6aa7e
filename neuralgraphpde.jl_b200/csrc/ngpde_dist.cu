// Multi-GPU behind the C ABI: the host-side node partitioner, a Morton ordering for
// arbitrarily numbered geometric graphs, an NCCL communicator bound at run time, and the per-RHS halo exchange.
//
// The reference is single-device (no collective call site under /root/reference); what it fixes is the semantics a
// partition must keep: `propagate` gathers x[s], x[t] per edge and reduces at t in stored edge order
// (/root/reference/src/layers.jl:111,326,416,534 through GraphNeuralNetworks [DEP]).  Hence owner-computes by destination
// with edges kept in their original relative order (forward bit-identical to one GPU), halo rows in ascending global id,
// and returned halo cotangents added per row in fixed peer order (deterministic backward).
//
// The plan arrays are produced by the same rules as the numpy statement in neuralgraphpde.jl_b200/partition.py
// (`partition_nodes`), which the CPU tests compare them with element by element.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "ngpde_common.cuh"

struct ngpde_partition {
  int world = 1, rank = 0;
  int64_t num_nodes = 0, lo = 0, hi = 0;
  std::vector<int64_t> arr[12];
};

namespace ngpde {
namespace {

template <class T>
inline int64_t idx_at(const void* p, int64_t i, int base) { return (int64_t) static_cast<const T*>(p)[i] - base; }

// contiguous node ranges: equal node counts, or equal (in-edges + 1) weight with cut positions found by a lower-bound
// search on the prefix sum (partition.balanced_bounds)
std::vector<int64_t> balanced_bounds(const std::vector<int64_t>& t, int64_t N, int world, bool by_edges) {
  std::vector<int64_t> b(world + 1);
  if (!by_edges || t.empty()) {
    for (int r = 0; r <= world; ++r) b[r] = (int64_t)r * N / world;
    return b;
  }
  std::vector<int64_t> cum(N + 1, 0);
  for (int64_t v : t) ++cum[v + 1];
  for (int64_t i = 0; i < N; ++i) cum[i + 1] += cum[i] + 1;  // + 1 per node: edgeless stretches are still spread out
  b[0] = 0;
  b[world] = N;
  for (int r = 1; r < world; ++r) {
    const int64_t target = (int64_t)r * cum[N] / world;
    b[r] = std::lower_bound(cum.begin(), cum.end(), target) - cum.begin();
  }
  for (int r = 1; r <= world; ++r) b[r] = std::max(b[r], b[r - 1]);
  return b;
}

// sorted distinct remote sources of the edges whose target lies in [lo, hi)
std::vector<int64_t> halo_of(const std::vector<int64_t>& s, const std::vector<int64_t>& t, int64_t lo, int64_t hi) {
  std::vector<int64_t> h;
  for (size_t e = 0; e < s.size(); ++e)
    if (t[e] >= lo && t[e] < hi && (s[e] < lo || s[e] >= hi)) h.push_back(s[e]);
  std::sort(h.begin(), h.end());
  h.erase(std::unique(h.begin(), h.end()), h.end());
  return h;
}

// ---- NCCL, bound at run time ----
struct UniqueId { char internal[NGPDE_UNIQUE_ID_BYTES]; };  // ncclUniqueId is passed BY VALUE to ncclCommInitRank
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
};
constexpr int kNcclFloat32 = 7, kNcclSum = 0;  // nccl.h: ncclFloat32 = 7, ncclSum = 0

NcclApi g_nccl;

int load_nccl(const char* path) {
  if (g_nccl.lib) return NGPDE_OK;
  const char* name = (path && *path) ? path : "libnccl.so.2";
  void* h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("cannot load NCCL (%s): %s", name, dlerror());
    return NGPDE_ERR_UNSUPPORTED;
  }
  NcclApi a;
  a.lib = h;
#define NGPDE_SYM(field, sym)                                              \
  *reinterpret_cast<void**>(&a.field) = dlsym(h, sym);                     \
  if (!a.field) {                                                          \
    set_error("NCCL symbol %s not found in %s", sym, name);                \
    return NGPDE_ERR_UNSUPPORTED;                                          \
  }
  NGPDE_SYM(GetUniqueId, "ncclGetUniqueId")
  NGPDE_SYM(CommInitRank, "ncclCommInitRank")
  NGPDE_SYM(CommDestroy, "ncclCommDestroy")
  NGPDE_SYM(GetErrorString, "ncclGetErrorString")
  NGPDE_SYM(GroupStart, "ncclGroupStart")
  NGPDE_SYM(GroupEnd, "ncclGroupEnd")
  NGPDE_SYM(Send, "ncclSend")
  NGPDE_SYM(Recv, "ncclRecv")
  NGPDE_SYM(AllReduce, "ncclAllReduce")
#undef NGPDE_SYM
  g_nccl = a;
  return NGPDE_OK;
}

#define NGPDE_NCCL_TRY(expr)                                                                   \
  do {                                                                                         \
    int r__ = (expr);                                                                          \
    if (r__ != 0) {                                                                            \
      ::ngpde::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(r__)); \
      return NGPDE_ERR_CUDA;                                                                   \
    }                                                                                          \
  } while (0)

}  // namespace
}  // namespace ngpde

struct ngpde_comm {
  void* comm = nullptr;
  int world = 1, rank = 0;
  bool owned = true;
};

struct ngpde_halo {
  ngpde_comm* comm = nullptr;
  int world = 1, rank = 0;
  int64_t n_owned = 0, n_halo = 0, n_send = 0, n_segs = 0;
  std::vector<int64_t> send_counts, recv_counts;
  int* d_send_rows = nullptr;
  int* d_seg_rows = nullptr;
  int* d_seg_ptr = nullptr;
  int* d_seg_pos = nullptr;
  float* buf = nullptr;  // pack buffer (forward) / returned cotangents (backward): n_send rows
  size_t buf_floats = 0;
  float* hbuf = nullptr;  // contiguous copy of the halo cotangent rows when dx_local's tail is not directly usable
};

using namespace ngpde;

extern "C" int ngpde_partition_create(ngpde_partition_t* out, int64_t num_nodes, int64_t num_edges, const void* src,
                                      const void* dst, int32_t index_dtype, int32_t index_base, int32_t world, int32_t rank,
                                      int32_t by_edges, const int64_t* bounds_in) {
  NGPDE_REQUIRE(out != nullptr, "partition: out is NULL");
  NGPDE_REQUIRE(world >= 1 && rank >= 0 && rank < world, "partition: rank %d of world %d", rank, world);
  NGPDE_REQUIRE(num_nodes >= 0 && num_edges >= 0 && (num_edges == 0 || (src && dst)), "partition: bad graph arguments");
  NGPDE_REQUIRE(index_dtype == NGPDE_IDX_I32 || index_dtype == NGPDE_IDX_I64, "partition: unknown index dtype %d", index_dtype);
  std::vector<int64_t> s(num_edges), t(num_edges);
  for (int64_t e = 0; e < num_edges; ++e) {
    s[e] = index_dtype == NGPDE_IDX_I64 ? idx_at<int64_t>(src, e, index_base) : idx_at<int32_t>(src, e, index_base);
    t[e] = index_dtype == NGPDE_IDX_I64 ? idx_at<int64_t>(dst, e, index_base) : idx_at<int32_t>(dst, e, index_base);
    NGPDE_REQUIRE(s[e] >= 0 && s[e] < num_nodes && t[e] >= 0 && t[e] < num_nodes, "partition: edge %lld (%lld -> %lld) is out of range",
                  (long long)e, (long long)s[e], (long long)t[e]);
  }
  std::vector<int64_t> b;
  if (bounds_in) {
    b.assign(bounds_in, bounds_in + world + 1);
    bool ok = b[0] == 0 && b[world] == num_nodes;
    for (int r = 0; r < world; ++r) ok = ok && b[r + 1] >= b[r];
    NGPDE_REQUIRE(ok, "partition: bounds must be a non-decreasing [world+1] vector from 0 to num_nodes");
  } else {
    b = balanced_bounds(t, num_nodes, world, by_edges != 0);
  }
  auto* p = new ngpde_partition();
  p->world = world;
  p->rank = rank;
  p->num_nodes = num_nodes;
  const int64_t lo = b[rank], hi = b[rank + 1];
  p->lo = lo;
  p->hi = hi;
  std::vector<int64_t> halo = halo_of(s, t, lo, hi);
  std::vector<int64_t> recv(world, 0), sendc(world, 0), peer_off(world, 0);
  for (int64_t h : halo) {
    const int owner = (int)(std::upper_bound(b.begin(), b.end(), h) - b.begin()) - 1;
    ++recv[owner];
  }
  std::vector<int64_t> s_local, t_local, eid;
  for (int64_t e = 0; e < num_edges; ++e) {
    if (t[e] < lo || t[e] >= hi) continue;
    eid.push_back(e);
    t_local.push_back(t[e] - lo);
    if (s[e] >= lo && s[e] < hi) s_local.push_back(s[e] - lo);
    else s_local.push_back((hi - lo) + (std::lower_bound(halo.begin(), halo.end(), s[e]) - halo.begin()));
  }
  // what every peer imports from me: the same computation from the peer's point of view
  std::vector<int64_t> send_local;
  for (int q = 0; q < world; ++q) {
    if (q == rank || hi == lo) continue;
    const std::vector<int64_t> hq = halo_of(s, t, b[q], b[q + 1]);
    const auto first = std::lower_bound(hq.begin(), hq.end(), lo), last = std::lower_bound(hq.begin(), hq.end(), hi);
    for (auto it = first; it != last; ++it) send_local.push_back(*it - lo);
    sendc[q] = last - first;
    peer_off[q] = first - hq.begin();  // q's halo rows owned by ranks below me come first
  }
  // segments: positions of the receive buffer grouped by owned row, stable (= ascending peer, fixed order)
  std::vector<int64_t> order(send_local.size());
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t c) { return send_local[a] < send_local[c]; });
  std::vector<int64_t> seg_rows, seg_ptr;
  for (size_t i = 0; i < order.size(); ++i) {
    if (i == 0 || send_local[order[i]] != send_local[order[i - 1]]) {
      seg_rows.push_back(send_local[order[i]]);
      seg_ptr.push_back((int64_t)i);
    }
  }
  seg_ptr.push_back((int64_t)order.size());
  if (order.empty()) seg_ptr.assign(1, 0);
  p->arr[NGPDE_PA_BOUNDS] = b;
  p->arr[NGPDE_PA_HALO_GLOBAL] = halo;
  p->arr[NGPDE_PA_RECV_COUNTS] = recv;
  p->arr[NGPDE_PA_SEND_COUNTS] = sendc;
  p->arr[NGPDE_PA_SEND_LOCAL] = send_local;
  p->arr[NGPDE_PA_S_LOCAL] = s_local;
  p->arr[NGPDE_PA_T_LOCAL] = t_local;
  p->arr[NGPDE_PA_EDGE_IDS] = eid;
  p->arr[NGPDE_PA_SEG_ROWS] = seg_rows;
  p->arr[NGPDE_PA_SEG_PTR] = seg_ptr;
  p->arr[NGPDE_PA_SEG_POS] = order;
  p->arr[NGPDE_PA_PEER_RECV_OFFSET] = peer_off;
  *out = p;
  return NGPDE_OK;
}

extern "C" int ngpde_partition_destroy(ngpde_partition_t p) {
  delete p;
  return NGPDE_OK;
}

extern "C" int ngpde_partition_array(ngpde_partition_t p, int32_t which, const int64_t** host_ptr, int64_t* len) {
  NGPDE_REQUIRE(p && host_ptr && len, "partition_array: null argument");
  NGPDE_REQUIRE(which >= 0 && which < 12, "partition_array: unknown array %d", which);
  *host_ptr = p->arr[which].data();
  *len = (int64_t)p->arr[which].size();
  return NGPDE_OK;
}

// Morton (Z-order) curve: coordinates normalised to [0, 1] over their bounding box, quantised to 21 bits per axis, bits
// interleaved (axis 0 most significant within a triple/pair); ties broken by node id (stable sort).
extern "C" int ngpde_morton_order(const float* pos, int64_t num_nodes, int32_t dim, int64_t* order) {
  NGPDE_REQUIRE(pos && order && num_nodes >= 0 && dim >= 1 && dim <= 3, "morton_order: bad argument");
  float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0};
  for (int a = 0; a < dim; ++a) {
    mn[a] = INFINITY;
    mx[a] = -INFINITY;
  }
  for (int64_t i = 0; i < num_nodes; ++i)
    for (int a = 0; a < dim; ++a) {
      mn[a] = std::min(mn[a], pos[i * dim + a]);
      mx[a] = std::max(mx[a], pos[i * dim + a]);
    }
  const int bits = 21;
  std::vector<uint64_t> code(num_nodes);
  for (int64_t i = 0; i < num_nodes; ++i) {
    uint64_t q[3] = {0, 0, 0};
    for (int a = 0; a < dim; ++a) {
      const double span = (double)mx[a] - (double)mn[a];
      const double u = span > 0 ? ((double)pos[i * dim + a] - (double)mn[a]) / span : 0.0;
      q[a] = (uint64_t)std::min<double>((double)((1u << bits) - 1), std::floor(u * (double)(1u << bits)));
    }
    uint64_t c = 0;
    for (int bit = bits - 1; bit >= 0; --bit)
      for (int a = 0; a < dim; ++a) c = (c << 1) | ((q[a] >> bit) & 1u);
    code[i] = c;
  }
  std::iota(order, order + num_nodes, (int64_t)0);
  std::stable_sort(order, order + num_nodes, [&](int64_t a, int64_t b) { return code[a] < code[b]; });
  return NGPDE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// communicator
// ---------------------------------------------------------------------------------------------------------------

extern "C" int ngpde_comm_unique_id(void* id_out, const char* libnccl_path) {
  NGPDE_REQUIRE(id_out != nullptr, "comm_unique_id: null argument");
  if (int rc = load_nccl(libnccl_path)) return rc;
  NGPDE_NCCL_TRY(g_nccl.GetUniqueId(id_out));
  return NGPDE_OK;
}

extern "C" int ngpde_comm_init(ngpde_comm_t* out, const void* unique_id, int32_t world, int32_t rank, const char* libnccl_path) {
  NGPDE_REQUIRE(out && unique_id && world >= 1 && rank >= 0 && rank < world, "comm_init: bad argument");
  if (int rc = load_nccl(libnccl_path)) return rc;
  UniqueId id;
  std::memcpy(id.internal, unique_id, NGPDE_UNIQUE_ID_BYTES);
  void* c = nullptr;
  NGPDE_NCCL_TRY(g_nccl.CommInitRank(&c, world, id, rank));
  auto* h = new ngpde_comm();
  h->comm = c;
  h->world = world;
  h->rank = rank;
  h->owned = true;
  *out = h;
  return NGPDE_OK;
}

extern "C" int ngpde_comm_adopt(ngpde_comm_t* out, void* nccl_comm, int32_t world, int32_t rank, const char* libnccl_path) {
  NGPDE_REQUIRE(out && nccl_comm && world >= 1 && rank >= 0 && rank < world, "comm_adopt: bad argument");
  if (int rc = load_nccl(libnccl_path)) return rc;
  auto* h = new ngpde_comm();
  h->comm = nccl_comm;
  h->world = world;
  h->rank = rank;
  h->owned = false;
  *out = h;
  return NGPDE_OK;
}

extern "C" int ngpde_comm_destroy(ngpde_comm_t c) {
  if (!c) return NGPDE_OK;
  if (c->owned && c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
  return NGPDE_OK;
}

extern "C" int ngpde_allreduce_sum(ngpde_comm_t c, float* buf, int64_t n, void* stream) {
  NGPDE_REQUIRE(c && (n == 0 || buf) && n >= 0, "allreduce_sum: bad argument");
  if (n == 0 || c->world == 1) return NGPDE_OK;
  NGPDE_NCCL_TRY(g_nccl.AllReduce(buf, buf, (size_t)n, kNcclFloat32, kNcclSum, c->comm, static_cast<cudaStream_t>(stream)));
  return NGPDE_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// halo exchange
// ---------------------------------------------------------------------------------------------------------------

namespace {
int upload_i32(const std::vector<int64_t>& v, int** dptr, cudaStream_t st) {
  *dptr = nullptr;
  std::vector<int> h(v.begin(), v.end());
  NGPDE_CUDA_TRY(cudaMalloc(dptr, sizeof(int) * std::max<size_t>(h.size(), 1)));
  if (!h.empty()) NGPDE_CUDA_TRY(cudaMemcpyAsync(*dptr, h.data(), sizeof(int) * h.size(), cudaMemcpyHostToDevice, st));
  NGPDE_CUDA_TRY(cudaStreamSynchronize(st));  // `h` dies at scope end
  return NGPDE_OK;
}
int ensure_buf(ngpde_halo* h, size_t floats) {
  if (h->buf_floats >= floats) return NGPDE_OK;
  if (h->buf) cudaFree(h->buf);
  if (h->hbuf) cudaFree(h->hbuf);
  h->buf = h->hbuf = nullptr;
  h->buf_floats = 0;
  NGPDE_CUDA_TRY(cudaMalloc(&h->buf, sizeof(float) * std::max<size_t>(floats, 1)));
  h->buf_floats = floats;
  return NGPDE_OK;
}
}  // namespace

extern "C" int ngpde_halo_create(ngpde_halo_t* out, ngpde_partition_t plan, ngpde_comm_t comm, void* stream) {
  NGPDE_REQUIRE(out && plan, "halo_create: null argument");
  NGPDE_REQUIRE(plan->world == 1 || comm, "halo_create: a communicator is needed for world > 1");
  NGPDE_REQUIRE(!comm || (comm->world == plan->world && comm->rank == plan->rank), "halo_create: plan and communicator disagree on (world, rank)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto* h = new ngpde_halo();
  h->comm = comm;
  h->world = plan->world;
  h->rank = plan->rank;
  h->n_owned = plan->hi - plan->lo;
  h->n_halo = (int64_t)plan->arr[NGPDE_PA_HALO_GLOBAL].size();
  h->n_send = (int64_t)plan->arr[NGPDE_PA_SEND_LOCAL].size();
  h->n_segs = (int64_t)plan->arr[NGPDE_PA_SEG_ROWS].size();
  h->send_counts = plan->arr[NGPDE_PA_SEND_COUNTS];
  h->recv_counts = plan->arr[NGPDE_PA_RECV_COUNTS];
  int rc = upload_i32(plan->arr[NGPDE_PA_SEND_LOCAL], &h->d_send_rows, st);
  if (!rc) rc = upload_i32(plan->arr[NGPDE_PA_SEG_ROWS], &h->d_seg_rows, st);
  if (!rc) rc = upload_i32(plan->arr[NGPDE_PA_SEG_PTR], &h->d_seg_ptr, st);
  if (!rc) rc = upload_i32(plan->arr[NGPDE_PA_SEG_POS], &h->d_seg_pos, st);
  if (rc) {
    ngpde_halo_destroy(h);
    return rc;
  }
  *out = h;
  return NGPDE_OK;
}

extern "C" int ngpde_halo_destroy(ngpde_halo_t h) {
  if (!h) return NGPDE_OK;
  cudaFree(h->d_send_rows);
  cudaFree(h->d_seg_rows);
  cudaFree(h->d_seg_ptr);
  cudaFree(h->d_seg_pos);
  cudaFree(h->buf);
  cudaFree(h->hbuf);
  delete h;
  return NGPDE_OK;
}

extern "C" int ngpde_halo_forward(ngpde_halo_t h, const float* x_owned, int32_t d, float* x_local, void* stream) {
  NGPDE_REQUIRE(h && d > 0 && (h->n_owned == 0 || (x_owned && x_local)), "halo_forward: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h->n_owned > 0 && x_local != x_owned)
    NGPDE_CUDA_TRY(cudaMemcpyAsync(x_local, x_owned, sizeof(float) * (size_t)h->n_owned * d, cudaMemcpyDeviceToDevice, st));
  if (h->world == 1) return NGPDE_OK;
  if (int rc = ensure_buf(h, (size_t)h->n_send * d)) return rc;
  if (int rc = ngpde_rows_gather(x_owned, h->d_send_rows, h->n_send, d, h->buf, stream)) return rc;
  float* halo = x_local + (size_t)h->n_owned * d;
  NGPDE_NCCL_TRY(g_nccl.GroupStart());
  size_t so = 0, ro = 0;
  for (int p = 0; p < h->world; ++p) {
    const size_t ns = (size_t)h->send_counts[p] * d, nr = (size_t)h->recv_counts[p] * d;
    if (ns) NGPDE_NCCL_TRY(g_nccl.Send(h->buf + so, ns, kNcclFloat32, p, h->comm->comm, st));
    if (nr) NGPDE_NCCL_TRY(g_nccl.Recv(halo + ro, nr, kNcclFloat32, p, h->comm->comm, st));
    so += ns;
    ro += nr;
  }
  NGPDE_NCCL_TRY(g_nccl.GroupEnd());
  return NGPDE_OK;
}

extern "C" int ngpde_halo_backward(ngpde_halo_t h, const float* dx_local, int32_t d, float* dx_owned, void* stream) {
  NGPDE_REQUIRE(h && d > 0 && (h->n_owned == 0 || (dx_local && dx_owned)), "halo_backward: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h->n_owned > 0 && dx_owned != dx_local)
    NGPDE_CUDA_TRY(cudaMemcpyAsync(dx_owned, dx_local, sizeof(float) * (size_t)h->n_owned * d, cudaMemcpyDeviceToDevice, st));
  if (h->world == 1) return NGPDE_OK;
  if (int rc = ensure_buf(h, (size_t)h->n_send * d)) return rc;
  const float* halo = dx_local + (size_t)h->n_owned * d;
  // the transpose of the forward exchange: what I received from p goes back to p, what I sent comes home
  NGPDE_NCCL_TRY(g_nccl.GroupStart());
  size_t so = 0, ro = 0;
  for (int p = 0; p < h->world; ++p) {
    const size_t ns = (size_t)h->recv_counts[p] * d, nr = (size_t)h->send_counts[p] * d;
    if (ns) NGPDE_NCCL_TRY(g_nccl.Send(halo + so, ns, kNcclFloat32, p, h->comm->comm, st));
    if (nr) NGPDE_NCCL_TRY(g_nccl.Recv(h->buf + ro, nr, kNcclFloat32, p, h->comm->comm, st));
    so += ns;
    ro += nr;
  }
  NGPDE_NCCL_TRY(g_nccl.GroupEnd());
  return ngpde_rows_segment_add(dx_owned, h->buf, h->d_seg_rows, h->d_seg_ptr, h->d_seg_pos, h->n_segs, d, stream);
}
