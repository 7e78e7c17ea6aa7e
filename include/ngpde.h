/* libngpde -- C ABI of the B200-native message-passing hot path of NeuralGraphPDE.jl.
 *
 * The reference has no FFI of its own: its hot path is the Julia call chain
 *     layer(x, ps, st)  ->  GraphNeuralNetworks.propagate  ->  NNlib.gather / Lux.Dense / NNlib.scatter
 * (/root/reference/src/layers.jl:98-112, 200-239, 312-332, 390-422, 509-547).  This header is the boundary a
 * Julia `ccall` shim or the Python/ctypes mirror (neuralgraphpde.jl_b200/) binds instead of
 * that chain.  Plain pointers and sizes only; every tensor is caller-owned DEVICE memory, float32, laid out
 * as Julia stores it: a feature matrix `(D, items)` column-major == C `[items][D]` row-major; a Lux Dense
 * weight `(out, in)` column-major == C `[in][out]`; parameters of one MLP are the flat ComponentArray
 * segment `weight_1, bias_1, weight_2, bias_2, ...`.
 *
 * Every function returns 0 on success or a negative NGPDE_ERR_* code; `ngpde_last_error()` returns a
 * thread-local message.  Calls are asynchronous on the caller-supplied `cudaStream_t` (passed as void*)
 * except graph construction, which synchronises that stream once.  No C++ exceptions cross the boundary.
 */
#ifndef NGPDE_H
#define NGPDE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NGPDE_VERSION 200

/* status codes */
#define NGPDE_OK 0
#define NGPDE_ERR_INVALID (-1) /* bad argument (mirrors the reference's @assert failures, layers.jl:204-207) */
#define NGPDE_ERR_CUDA (-2)    /* a CUDA runtime call failed */
#define NGPDE_ERR_WORKSPACE (-3) /* workspace too small */
#define NGPDE_ERR_UNSUPPORTED (-4)

/* activation functions of a Lux Dense layer (NNlib names) */
enum {
  NGPDE_ACT_IDENTITY = 0,
  NGPDE_ACT_RELU = 1,
  NGPDE_ACT_TANH = 2,
  NGPDE_ACT_SIGMOID = 3,
  NGPDE_ACT_SWISH = 4,
  NGPDE_ACT_GELU = 5, /* NNlib 0.8 gelu: tanh approximation */
  NGPDE_ACT_SOFTPLUS = 6,
  NGPDE_ACT_ELU = 7,
  NGPDE_ACT_LEAKYRELU = 8 /* slope 0.01 */
};

/* aggregation operators of propagate(..., aggr) -- NNlib.scatter semantics: sequential in stored edge
 * order per destination; identity 0 / 0 / -Inf / +Inf / 1 for isolated nodes; mean = sum / float(count).
 * NGPDE_AGGR_PROD is `aggr = *` (/root/reference/src/layers.jl:49,257,348,441); its pullback is NNlib's
 * product of the OTHER messages of the destination (never a division by the message: zeros are exact).
 * It runs on the FFMA kernels only; the persistent ODE kernels and the factored GNOConv take + and mean. */
enum { NGPDE_AGGR_SUM = 0, NGPDE_AGGR_MEAN = 1, NGPDE_AGGR_MAX = 2, NGPDE_AGGR_MIN = 3, NGPDE_AGGR_PROD = 4 };

/* layer families (/root/reference/src/layers.jl) */
enum {
  NGPDE_EXPLICIT_EDGE_CONV = 0, /* :84-112  */
  NGPDE_VMH_CONV = 1,           /* :295-332 */
  NGPDE_MPPDE_CONV = 2,         /* :377-422 */
  NGPDE_GNO_CONV = 3            /* :485-547 */
};

/* index arrays handed to ngpde_graph_create */
enum { NGPDE_IDX_I32 = 0, NGPDE_IDX_I64 = 1 };

/* integer arrays a graph handle exposes (ngpde_graph_array) -- part of the bit-exact index contract */
enum {
  NGPDE_GA_ROWPTR = 0,   /* int32 [N+1]  dst-sorted CSR row pointers                                  */
  NGPDE_GA_SRC = 1,      /* int32 [E]    source of the k-th edge in CSR order                          */
  NGPDE_GA_DST = 2,      /* int32 [E]    destination of the k-th edge in CSR order                     */
  NGPDE_GA_PERM = 3,     /* int32 [E]    original COO position of the k-th CSR edge (stable sort)      */
  NGPDE_GA_TPTR = 4,     /* int32 [N+1]  src-sorted transpose pointers                                 */
  NGPDE_GA_TPOS = 5,     /* int32 [E]    CSR position of the k-th edge in src-sorted order (stable)    */
  NGPDE_GA_UNITS32 = 6,  /* int32 [U+1]  work-unit node boundaries for 32-edge tiles                   */
  NGPDE_GA_UNITS64 = 7,  /*              ... 64-edge tiles                                             */
  NGPDE_GA_UNITS128 = 8, /*              ... 128-edge tiles                                            */
  NGPDE_GA_GCN_COLPTR = 9,  /* int32 [N+1]  merged (dst, ascending src) adjacency, see ngpde_gcn_*     */
  NGPDE_GA_GCN_ROWVAL = 10, /* int32 [nnz]                                                             */
  NGPDE_GA_GCN_TPTR = 11,   /* int32 [N+1]  its transpose (src, ascending dst)                         */
  NGPDE_GA_GCN_TPOS = 12    /* int32 [nnz]  merged-entry index of the k-th transposed entry            */
};

#define NGPDE_MAX_LAYERS 8

/* A Lux `Chain` of `Dense` layers (or one bare `Dense`): layer l maps dims[l] -> dims[l+1]. */
typedef struct {
  int32_t n_layers;
  int32_t dims[NGPDE_MAX_LAYERS + 1];
  int32_t act[NGPDE_MAX_LAYERS];
  int32_t has_bias[NGPDE_MAX_LAYERS];
} ngpde_mlp;

/* One message-passing layer call.  Which fields are read depends on `family`:
 *
 *  EXPLICIT_EDGE_CONV  m_k = phi([h_t; h_s; pos_s - pos_t]);           y = aggr_t(m)            (layers.jl:98-112)
 *  VMH_CONV            m_k = phi([h_t; h_s - h_t; pos_s - pos_t]);     y = node([x; aggr_t(m)]) (layers.jl:312-332)
 *      h = [x (dx cols, differentiable); snode[:, 0:dhs]],  pos = snode[:, dhs:dhs+dpos]
 *  MPPDE_CONV          m_k = phi([x_t; x_s; S_t - S_s; e_k; theta_g(k)]); y = node([x; aggr(m); theta_g(i)])
 *      S = snode (ds = dhs + dpos cols), theta indexed by (original edge position) / (E / G)      (layers.jl:390-422)
 *  GNO_CONV            W_k = reshape(phi([S_t; S_s; e_k]), out, in);   m_k = W_k x_s;
 *                      y = act(W_lin x + aggr_t(m) + b)   -- `node` is the 1-layer `linear` Dense (layers.jl:509-547)
 */
typedef struct {
  int32_t family;
  int32_t aggr;
  int32_t dx;     /* width of the differentiable node state x                                  */
  int32_t dhs;    /* static columns of snode that ride with h (EdgeConv / VMH), else part of S */
  int32_t dpos;   /* position columns of snode (EdgeConv / VMH), else part of S                */
  int32_t de;     /* width of edata                                                            */
  int32_t dtheta; /* width of theta                                                            */
  int32_t gno_in, gno_out;
  ngpde_mlp phi;  /* edge MLP                                                                  */
  ngpde_mlp node; /* gamma / psi / linear; n_layers == 0 for EXPLICIT_EDGE_CONV                */
} ngpde_conv_desc;

typedef struct {
  const float* x;          /* [N][dx]                                   */
  const float* snode;      /* [N][dhs + dpos] static node data, or NULL */
  const float* edata;      /* [E][de] in ORIGINAL edge order, or NULL   */
  const float* theta;      /* [G][dtheta], or NULL                      */
  const float* phi_params; /* flat parameters of phi                    */
  const float* node_params;/* flat parameters of gamma / psi / linear   */
  float* mbar;             /* [N][dm] aggregated messages (written by forward, read by backward) */
  float* y;                /* [N][dy] layer output (for EXPLICIT_EDGE_CONV this is mbar itself)  */
  /* backward only */
  const float* dy;         /* [N][dy] cotangent of y                    */
  float* dx;               /* [N][dx] cotangent of x (overwritten)      */
  float* dphi_params;      /* flat, overwritten                         */
  float* dnode_params;     /* flat, overwritten                         */
  /* optional (may be NULL; forward and backward): ngpde_conv_state_bytes(g, desc) bytes of caller-owned device memory,
   * 16-byte aligned, in which the forward leaves what the backward would otherwise recompute (the hoisted first-layer
   * projections).  The backward may be given it only after the matching forward call -- same graph,
   * descriptor, x, parameters and static data -- and before anything else writes to it. */
  void* state;
} ngpde_conv_io;

typedef struct ngpde_graph* ngpde_graph_t;

/* ---- library ---- */
int ngpde_version(void);
const char* ngpde_last_error(void);
/* Process-wide switches (benchmarks and tests only).  NGPDE_OPT_TENSOR_CORES: 1 (default) runs MLPs whose layers are at
 * most 64 wide on the tcgen05 tensor-core kernels (3xTF32, fp32-accurate); 0 forces the FP32-FFMA kernels everywhere. */
enum { NGPDE_OPT_TENSOR_CORES = 0, NGPDE_OPT_GNO_FACTORED = 1, NGPDE_OPT_DEBUG_SKIP = 2 /* developer aid: phase timing */,
       NGPDE_OPT_HOIST = 3, NGPDE_OPT_LAYERED = 4, NGPDE_OPT_GNO_LAYERED = 5 };
/* NGPDE_OPT_GNO_LAYERED: 1 (default) runs the factored GNOConv (hidden width == in_chs in {32, 64}, >= 2048 edges) with
 * phi's hidden layers on the tcgen05 GEMM and the per-destination products in a warp-per-node kernel
 * (csrc/ngpde_gno_node.cuh); 0 keeps the fused FFMA edge kernels. */
/* NGPDE_OPT_LAYERED: 1 (default) evaluates layers whose MLPs are wider than the fused tensor-core kernels take (an output
 * wider than 64 columns; + / mean aggregation; at least 8192 edges) one Dense layer at a time on the tcgen05 GEMM
 * (csrc/ngpde_layered.cuh); 0 keeps them on the fused FP32-FFMA kernels. */
/* NGPDE_OPT_GNO_FACTORED: 1 (default) evaluates GNOConv whose phi ends in an affine layer, with aggr = + or mean, in
 * factored form (per-destination outer-product sums + dense GEMMs; csrc/ngpde_gno.cuh); 0 forces the per-edge contraction. */
int ngpde_set_option(int32_t option, int32_t value);

/* ---- graph handle: replaces GNNGraph's per-call gather/scatter index use and GCNConv's per-call
 * add_self_loops / degree / adjacency_matrix (layers.jl:210-225).  Builds, on the device: the stable
 * dst-sorted CSR, the edge permutation, in-degrees, the src-sorted transpose and the work-unit lists.
 * `src`/`dst` are E indices (host or device memory, int32 or int64, `index_base` 0 or 1) in the stored
 * COO order.  `num_graphs` equal-sized graphs are assumed laid out contiguously (layers.jl:410,418). ---- */
int ngpde_graph_create(ngpde_graph_t* out, int64_t num_nodes, int64_t num_edges, const void* src, const void* dst,
                       int32_t index_dtype, int32_t index_base, int32_t indices_on_device, int64_t num_graphs,
                       void* stream);
int ngpde_graph_destroy(ngpde_graph_t g);
int ngpde_graph_array(ngpde_graph_t g, int32_t which, int32_t with_self_loops, const void** device_ptr, int64_t* len);
/* copy one of those arrays into caller-owned device memory holding at least `capacity` int32 */
int ngpde_graph_array_copy(ngpde_graph_t g, int32_t which, int32_t with_self_loops, int32_t* dst_device,
                           int64_t capacity, void* stream);
int64_t ngpde_graph_num_nodes(ngpde_graph_t g);
int64_t ngpde_graph_num_edges(ngpde_graph_t g);

/* ---- bare propagate(copy_xj | e_mul_xj, g, aggr; xj = x [, e = w]) on the scatter path: ordered,
 * atomic-free aggregation (call sites layers.jl:228-232, 656).  w may be NULL; it is [E] in ORIGINAL order. ---- */
int ngpde_aggregate(ngpde_graph_t g, int32_t aggr, const float* x, int32_t d, const float* w, float* out, void* stream);

/* ---- MLP message-passing layers.  Both calls need a caller-owned device workspace of at least
 * ngpde_conv_workspace_bytes(g, desc, backward) bytes, 256-byte aligned (forward: prepared weight images of the
 * tensor-core path; backward: transposed weights, per-edge source gradients, per-CTA parameter-gradient partials). ---- */
size_t ngpde_conv_workspace_bytes(ngpde_graph_t g, const ngpde_conv_desc* desc, int32_t backward);
/* bytes of the optional ngpde_conv_io.state buffer (0: this descriptor keeps nothing between forward and backward) */
size_t ngpde_conv_state_bytes(ngpde_graph_t g, const ngpde_conv_desc* desc);
int ngpde_conv_forward(ngpde_graph_t g, const ngpde_conv_desc* desc, const ngpde_conv_io* io, void* workspace,
                       size_t workspace_bytes, void* stream);
int ngpde_conv_backward(ngpde_graph_t g, const ngpde_conv_desc* desc, const ngpde_conv_io* io, void* workspace,
                        size_t workspace_bytes, void* stream);

/* per-family names (each checks desc->family and forwards to the two calls above) */
int ngpde_explicit_edge_conv_forward(ngpde_graph_t, const ngpde_conv_desc*, const ngpde_conv_io*, void*, size_t, void*);
int ngpde_explicit_edge_conv_backward(ngpde_graph_t, const ngpde_conv_desc*, const ngpde_conv_io*, void*, size_t, void*);
int ngpde_vmh_conv_forward(ngpde_graph_t, const ngpde_conv_desc*, const ngpde_conv_io*, void*, size_t, void*);
int ngpde_vmh_conv_backward(ngpde_graph_t, const ngpde_conv_desc*, const ngpde_conv_io*, void*, size_t, void*);
int ngpde_mppde_conv_forward(ngpde_graph_t, const ngpde_conv_desc*, const ngpde_conv_io*, void*, size_t, void*);
int ngpde_mppde_conv_backward(ngpde_graph_t, const ngpde_conv_desc*, const ngpde_conv_io*, void*, size_t, void*);
int ngpde_gno_conv_forward(ngpde_graph_t, const ngpde_conv_desc*, const ngpde_conv_io*, void*, size_t, void*);
int ngpde_gno_conv_backward(ngpde_graph_t, const ngpde_conv_desc*, const ngpde_conv_io*, void*, size_t, void*);

/* ---- GCNConv (layers.jl:200-239), CPU sparse-matmul semantics of GNN.jl: merged duplicate edges,
 * ascending-source accumulation order, multiply and add rounded separately.
 *   y = act(W (c .* A^T (c .* x)) + b),   c = 1/sqrt(in-degree),   W applied first when out < in.
 * edge_weight: NULL or [E] (original order; self-loop weights of 1 are appended internally, layers.jl:215).
 * graph_weight: NULL or the graph's own stored weights [E].  They scale the messages only when use_edge_weight != 0
 *   (w_mul_xj, layers.jl:230), but whenever present (and no explicit edge_weight is given) they weight the in-degree of
 *   the normaliser: `degree(g, T; dir=:in, edge_weight)` with edge_weight === nothing resolves to the stored weights in
 *   GNN.jl (layers.jl:224).
 * agg_buf: [N][min(in,out)] scratch kept for the backward; lin_buf: [N][min(in,out)] scratch. ---- */
typedef struct {
  int32_t in_chs, out_chs;
  int32_t act;
  int32_t has_bias;
  int32_t add_self_loops;
  int32_t use_edge_weight;
} ngpde_gcn_desc;

size_t ngpde_gcn_workspace_bytes(ngpde_graph_t g, const ngpde_gcn_desc* desc, int32_t backward);
int ngpde_gcn_conv_forward(ngpde_graph_t g, const ngpde_gcn_desc* desc, const float* x, const float* weight,
                           const float* bias, const float* edge_weight, const float* graph_weight, float* y,
                           void* workspace, size_t workspace_bytes, void* stream);
int ngpde_gcn_conv_backward(ngpde_graph_t g, const ngpde_gcn_desc* desc, const float* x, const float* weight,
                            const float* bias, const float* edge_weight, const float* graph_weight, const float* y,
                            const float* dy, float* dx, float* dweight, float* dbias, void* workspace,
                            size_t workspace_bytes, void* stream);

/* ---- fixed-step ODE glue (the immediate caller of the path):
 * out = u + sum_i coef[i] * k[i]  over n floats, nk <= 8 stage arrays; u == NULL stands for zeros.  ---- */
int ngpde_axpy_stages(float* out, const float* u, const float* const* k, const float* coef, int32_t nk, int64_t n,
                      void* stream);

/* ---- multi-GPU: halo exchange of a node-partitioned graph.  The reference is single-device
 * (no collective call site exists under /root/reference); a partitioned `propagate` (call sites layers.jl:111,326,416,534)
 * needs, per layer call, the boundary rows of x packed for every peer and -- in the pullback -- the returned halo
 * cotangents added into the owner's rows.  The transfer itself is NCCL (all-to-all over NVLink) in the host layer, or
 * direct peer stores (ngpde_rows_put) when the destination buffers are peer-mapped.
 *   rows_gather       out[i][:] = x[rows[i]][:]                                  i < n_rows
 *   rows_put          row i goes to peer p = the slot with peer_ptr[p] <= i < peer_ptr[p+1], at
 *                     peer_dst[p][(i - peer_ptr[p])][:]  (peer_ptr, peer_dst: DEVICE arrays; peer_dst[p] may point into
 *                     another GPU's memory)
 *   rows_segment_add  dst[seg_rows[u]][:] += sum over q in [seg_ptr[u], seg_ptr[u+1]) of src[seg_pos[q]][:], q ascending:
 *                     deterministic, atomic-free; seg_rows must be distinct. ---- */
int ngpde_rows_gather(const float* x, const int32_t* rows, int64_t n_rows, int32_t d, float* out, void* stream);
int ngpde_rows_put(const float* x, const int32_t* rows, const int64_t* peer_ptr, float* const* peer_dst, int32_t n_peers,
                   int64_t n_rows, int32_t d, void* stream);
int ngpde_rows_segment_add(float* dst, const float* src, const int32_t* seg_rows, const int32_t* seg_ptr,
                           const int32_t* seg_pos, int64_t n_segs, int32_t d, void* stream);
/* One-shot all-reduce of the flat parameter gradient over peer-mapped buffers (NVLink / NVSwitch): out[i] = sum over
 * r = 0..world-1, in that order on EVERY rank, of peer_bufs[r][i] -- deterministic and bit-identical across ranks, one
 * kernel, no NCCL call.  peer_bufs is a DEVICE array of `world` pointers (16-byte aligned buffers of n floats: this rank's
 * own and its peers' symmetric allocations).  The caller brackets the call with a cross-GPU barrier on both sides (all
 * ranks have written before; nobody overwrites its buffer until all have read). */
int ngpde_peer_allreduce_sum(const float* const* peer_bufs, int32_t world, float* out, int64_t n, void* stream);

/* ---- persistent fixed-step Runge-Kutta integrator for du/dt = ExplicitEdgeConv(u) on a small graph (SURVEY.md section 8f-1;
 * the reference's `solve(prob, Tsit5(); adaptive = false, dt)` loop, docs/src/tutorials/graph_node.md:53-66): ONE kernel
 * launch (a thread-block cluster, one cluster barrier per right-hand side) integrates `n_steps` steps; a second kernel is its
 * discrete adjoint.  Limits: family EXPLICIT_EDGE_CONV, aggr + / mean, dx <= 4, phi input <= 16, phi layers <= 32 wide,
 * <= 4 layers, <= 1024 parameters, activations whose derivative is a function of the output, 8 <= N <= 65536.
 *   tableau   explicit: a[i][j] (j < i) and b[i]; stage i input = u + dt sum_j a[i][j] k_j; u_next = u + dt sum_i b[i] k_i
 *   forward   u [N][dx] in/out; traj [n_steps][n_stages][N][dx] receives every stage input (the adjoint's checkpoint)
 *   adjoint   lam [N][dx]: in dL/du(T), out dL/du(0); dphi_params: dL/dparams (overwritten); deterministic (no atomics) ---- */
typedef struct {
  int32_t n_stages;
  float a[8][8];
  float b[8];
} ngpde_rk_tableau;
size_t ngpde_edgeconv_ode_workspace_bytes(ngpde_graph_t g, const ngpde_conv_desc* desc, const ngpde_rk_tableau* tab);
int ngpde_edgeconv_ode_forward(ngpde_graph_t g, const ngpde_conv_desc* desc, const ngpde_rk_tableau* tab, float dt, int32_t n_steps,
                               const float* phi_params, const float* snode, float* u, float* traj, void* workspace,
                               size_t workspace_bytes, void* stream);
int ngpde_edgeconv_ode_adjoint(ngpde_graph_t g, const ngpde_conv_desc* desc, const ngpde_rk_tableau* tab, float dt, int32_t n_steps,
                               const float* phi_params, const float* snode, const float* traj, float* lam, float* dphi_params,
                               void* workspace, size_t workspace_bytes, void* stream);

/* ---- multi-GPU, host side: the node partitioner.  Owner-computes by destination over contiguous
 * node ranges: rank r owns [bounds[r], bounds[r+1]) and every edge whose target it owns, in the original relative order
 * (so each owned row reduces the same messages in the same order as on one GPU: the forward is bit-identical); sources
 * owned elsewhere form the halo, appended after the owned rows in ascending global id.  `src`/`dst` are HOST arrays of
 * the full COO lists (every rank holds them at build time; nothing is communicated here).  bounds == NULL: balanced by
 * in-edge count + 1 per node (by_edges != 0) or by node count.  All arrays a plan exposes are int64 host arrays owned by
 * the plan; together with the graph handle's arrays they are the bit-exact index contract of the partitioned path. ---- */
typedef struct ngpde_partition* ngpde_partition_t;
enum {
  NGPDE_PA_BOUNDS = 0,       /* [world+1] global node ranges                                                      */
  NGPDE_PA_HALO_GLOBAL = 1,  /* [n_halo]  global ids of imported sources, ascending (= grouped by owner)          */
  NGPDE_PA_RECV_COUNTS = 2,  /* [world]   halo rows received from each peer                                       */
  NGPDE_PA_SEND_COUNTS = 3,  /* [world]   owned rows sent to each peer                                            */
  NGPDE_PA_SEND_LOCAL = 4,   /* [n_send]  owned-local row ids, grouped by peer, each group ascending              */
  NGPDE_PA_S_LOCAL = 5,      /* [E_local] local source ids (owned: id - lo; halo: n_owned + rank in HALO_GLOBAL)  */
  NGPDE_PA_T_LOCAL = 6,      /* [E_local] local target ids (always owned)                                         */
  NGPDE_PA_EDGE_IDS = 7,     /* [E_local] original COO positions, ascending                                       */
  NGPDE_PA_SEG_ROWS = 8,     /* [U]       distinct owned-local rows with at least one remote reader, ascending    */
  NGPDE_PA_SEG_PTR = 9,      /* [U+1]                                                                             */
  NGPDE_PA_SEG_POS = 10,     /* [n_send]  positions in the peer-major buffer of returned halo cotangents          */
  NGPDE_PA_PEER_RECV_OFFSET = 11 /* [world] row offset of my rows inside peer p's halo block (direct peer stores) */
};
int ngpde_partition_create(ngpde_partition_t* out, int64_t num_nodes, int64_t num_edges, const void* src, const void* dst,
                           int32_t index_dtype, int32_t index_base, int32_t world, int32_t rank, int32_t by_edges,
                           const int64_t* bounds);
int ngpde_partition_destroy(ngpde_partition_t p);
int ngpde_partition_array(ngpde_partition_t p, int32_t which, const int64_t** host_ptr, int64_t* len);
/* Morton (Z-order) permutation of the nodes from their coordinates pos [N][dim] (HOST, float32, dim = 1..3): order[k] = id of
 * the k-th node along the curve.  Relabelling the graph with it before ngpde_partition_create turns contiguous id ranges
 * into compact spatial blocks, so an arbitrarily numbered geometric graph gets an O(sqrt(N)) halo instead of O(N). */
int ngpde_morton_order(const float* pos, int64_t num_nodes, int32_t dim, int64_t* order);

/* ---- multi-GPU, device side: communicator + per-RHS halo exchange.  NCCL is bound at run time (dlopen of
 * `libnccl_path`, NULL = "libnccl.so.2": the copy the host framework already loaded -- torch's bundled one, NCCL.jl's
 * artifact), so libngpde itself has no link-time dependency on it.
 *   comm_unique_id     rank 0 creates the 128-byte id; the host layer ships it to the other ranks (any side channel)
 *   comm_init          ncclCommInitRank on the CURRENT device
 *   comm_adopt         wrap an existing ncclComm_t (not destroyed by comm_destroy)
 *   halo_create        uploads the plan's send / segment lists; owns a send buffer and a receive buffer sized on demand
 *   halo_forward       x_local[0:n_owned] = x_owned; boundary rows packed (ngpde_rows_gather) and exchanged with grouped
 *                      ncclSend/ncclRecv straight into x_local[n_owned:]            (forward of propagate's gather)
 *   halo_backward      dx_owned = dx_local[0:n_owned] + returned halo cotangents added per row in fixed peer order
 *                      (ngpde_rows_segment_add): deterministic                      (its pullback)
 *   allreduce_sum      in-place sum of the flat parameter gradient over ranks. ---- */
typedef struct ngpde_comm* ngpde_comm_t;
typedef struct ngpde_halo* ngpde_halo_t;
#define NGPDE_UNIQUE_ID_BYTES 128
int ngpde_comm_unique_id(void* id_out, const char* libnccl_path);
int ngpde_comm_init(ngpde_comm_t* out, const void* unique_id, int32_t world, int32_t rank, const char* libnccl_path);
int ngpde_comm_adopt(ngpde_comm_t* out, void* nccl_comm, int32_t world, int32_t rank, const char* libnccl_path);
int ngpde_comm_destroy(ngpde_comm_t c);
int ngpde_halo_create(ngpde_halo_t* out, ngpde_partition_t plan, ngpde_comm_t comm, void* stream);
int ngpde_halo_destroy(ngpde_halo_t h);
int ngpde_halo_forward(ngpde_halo_t h, const float* x_owned, int32_t d, float* x_local, void* stream);
int ngpde_halo_backward(ngpde_halo_t h, const float* dx_local, int32_t d, float* dx_owned, void* stream);
int ngpde_allreduce_sum(ngpde_comm_t c, float* buf, int64_t n, void* stream);

/* ---- the training step either side of the adjoint (docs/src/tutorials/graph_node.md:100-129,
 * VMH.md:97-109,140-143): loss value + cotangent, and the optimiser update over the flat parameter vector.
 *   adam_step   Optimisers.Adam:  m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; x -= m/(1-b1^t) / (sqrt(v/(1-b2^t)) + eps) * eta
 *               (beta1_t = b1^t, beta2_t = b2^t are carried by the host as Optimisers.jl does)
 *   rprop_step  Optimisers.Rprop: eta_i *= l+ (<= gamma_max) while g keeps its sign, *= l- (>= gamma_min) and g_prev = 0 when
 *               it flips; x -= eta_i sign(g_prev')
 *   mse_loss    loss = mean((yhat - y)^2) over n floats; dyhat (may be NULL) = 2 (yhat - y) / n
 *   logit_cross_entropy   yhat [n_rows][C]; mask_idx [n_masked] row ids (NULL: all rows); y [n_masked][C];
 *               loss = mean_j( -sum_c y[j][c] logsoftmax(yhat[mask[j]])[c] ); dyhat [n_rows][C] (may be NULL; zero outside the mask)
 * `loss` is a DEVICE scalar; reductions are two-stage in fixed order (deterministic). ---- */
int ngpde_adam_step(float* params, const float* grad, float* m, float* v, int64_t n, float eta, float beta1, float beta2,
                    float eps, float beta1_t, float beta2_t, void* stream);
int ngpde_rprop_step(float* params, const float* grad, float* g_prev, float* eta, int64_t n, float ell_minus, float ell_plus,
                     float gamma_min, float gamma_max, void* stream);
size_t ngpde_loss_workspace_bytes(void);
int ngpde_mse_loss(const float* yhat, const float* y, int64_t n, float* loss, float* dyhat, void* workspace,
                   size_t workspace_bytes, void* stream);
int ngpde_logit_cross_entropy(const float* yhat, int64_t n_rows, int32_t n_classes, const float* y, const int32_t* mask_idx,
                              int64_t n_masked, float* loss, float* dyhat, void* workspace, size_t workspace_bytes,
                              void* stream);

/* number of kernel nodes / of all nodes of a captured cudaGraph_t (benchmarks count a replayed step's launches with it) */
int ngpde_cuda_graph_kernel_nodes(void* cuda_graph, int64_t* n_kernels, int64_t* n_nodes);

/* ---- optional kernel timing: while enabled, the four fused kernels of the conv layers (edge/node phase, forward/
 * backward) are bracketed by CUDA events on the launching stream.  ngpde_profile_read synchronises those events, returns
 * the summed milliseconds and launch counts per slot (arrays of NGPDE_PROF_SLOTS) and clears the record.  Not
 * thread-safe; meant for benchmarks. ---- */
enum { NGPDE_PROF_FWD_EDGE = 0, NGPDE_PROF_FWD_NODE = 1, NGPDE_PROF_BWD_NODE = 2, NGPDE_PROF_BWD_EDGE = 3, NGPDE_PROF_SLOTS = 4 };
int ngpde_profile_enable(int32_t on);
/* developer aid: while a device buffer of 512 int64 is registered, CTA 0 of the tensor-core edge-phase backward writes
 * clock64() stamps of its phases for its first 16 tiles ([tile][32]); NULL switches it off. */
int ngpde_debug_buffer(void* device_int64_x512);
int ngpde_profile_read(double* total_ms, int64_t* launches);
/* which engine a layer's four fused kernels run on for this graph/descriptor: paths[NGPDE_PROF_*] = 1 for the tcgen05
 * tensor-core kernels (3xTF32), 0 for the FP32-FFMA engine, 2 for the factored GNOConv evaluation (FFMA edge kernel + dense
 * FP32 GEMMs), -1 when the layer has no such phase. */
int ngpde_conv_kernel_paths(ngpde_graph_t g, const ngpde_conv_desc* desc, int32_t* paths);

/* developer/test aid: the dense GEMM of the factored GNOConv evaluation (csrc/ngpde_gno.cuh) on its own.
 * C[M][N] = op(A) op(B) over K; a_kmajor: A stored [K][M] (else [M][K]); b_kmajor: B stored [K][N] (else [N][K]);
 * splits > 1 writes split-K slices C + s*M*ldc; deg_rowptr: divide row m by rowptr[m+1]-rowptr[m] (0 for empty rows);
 * engine 0 = FP32 FFMA, 1 = tcgen05 3xTF32. */
int ngpde_debug_gemm(const float* A, int32_t lda, int32_t a_kmajor, const float* B, int32_t ldb, int32_t b_kmajor, float* C,
                     int32_t ldc, int64_t M, int32_t N, int64_t K, int32_t splits, const int32_t* deg_rowptr,
                     int32_t engine, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NGPDE_H */
