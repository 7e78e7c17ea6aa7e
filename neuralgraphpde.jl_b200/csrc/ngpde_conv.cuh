// Kernel argument blocks of the fused message-passing kernels.
#pragma once
#include "ngpde_tile.cuh"

namespace ngpde {

// How one block of rows of the MLP input Z0 is assembled per column (edge or node):
enum { SEG_DST = 0, SEG_SRC = 1, SEG_SMD = 2 /* a[src]-a[dst] */, SEG_DMS = 3 /* a[dst]-a[src] */, SEG_EDGE = 4, SEG_GRAPH = 5 };
// which array a segment reads
enum { ARR_X = 0, ARR_S = 1, ARR_E = 2, ARR_T = 3, ARR_M = 4, ARR_COUNT = 5 };

struct Seg {
  int kind, arr, col, width, row;
};

struct MlpDev {
  int L;
  int dims[NGPDE_MAX_LAYERS + 1];
  int act[NGPDE_MAX_LAYERS];
  int w_off[NGPDE_MAX_LAYERS];   // float offset of weight [in][out] in the flat parameter segment
  int b_off[NGPDE_MAX_LAYERS];   // float offset of bias [out], or -1
  int n_params;
};

// Topology seen by one kernel.  In the node phase a "unit" is TE consecutive nodes and src = dst = perm = node id.
struct TileGraph {
  const int* rowptr;
  const int* src;
  const int* dst;
  const int* perm;
  const int* unit_ptr;
  int n_units;
  int N;
  int gdiv;  // items per graph: E/G (edge phase, indexed by ORIGINAL edge position) or N/G (node phase)
};

struct FwdArgs {
  TileGraph tg;
  const float* arr[ARR_COUNT];
  int ld[ARR_COUNT];
  int n_segs;
  Seg segs[8];
  MlpDev mlp;
  const float* params;
  int contract;  // GNO: 1 = last layer contracted with x[src] per edge; 2 = factored evaluation (ngpde_gno.cuh)
  int gin, gout;
  float* gno_S;  // contract == 2: [N][gno_Ka * gin] per-destination outer-product sums (workspace)
  int gno_Ka;    // rows of S_n: width of phi's last hidden layer (+1 when its last layer has a bias)
  int offZt;
  int aggr;
  int dout;             // rows of the result tile
  float* out;           // edge phase: mbar [N][dout]; node phase: y [N][dout]
  int out_ld;           // node phase, tensor-core kernels: leading dimension of `out` (0: dout)
  int skip_l0;          // tensor-core kernels: layer 0 is the identity of a hoisted first layer -- its activation is applied to the
                        // gathered input directly, no MMAs
  const float* addend;  // node phase (GNO): [N][dout] added before the last activation
  float* msg_out;       // edge phase, aggr = *: when set the per-edge messages are written here ([E][dout], CSR order) and nothing is
                        // aggregated (the backward's product-of-others pass reads them)
  int offA, offB, offW, offH;  // shared-memory float offsets
};

struct BwdArgs {
  TileGraph tg;
  const float* arr[ARR_COUNT];
  int ld[ARR_COUNT];
  int n_segs;
  Seg segs[8];
  MlpDev mlp;
  const float* params;
  const float* wt;  // transposed weights: layer l at wt + w_off[l], laid out [out][in]
  int contract, gin, gout;
  int aggr;
  int dout;
  const float* gout_ptr;  // edge phase: dmbar [N][dout]; node phase: dy [N][dout]
  const float* fwd_out;   // edge phase: mbar (max/min mask)
  const float* gedge;     // edge phase, aggr = *: per-edge message cotangents [E][dout] in CSR order (replace the gather of gout_ptr)
  const float* addend;    // node phase (GNO)
  float* dparams_partial; // [gridDim.x][n_params], zero-initialised
  float* dx_direct;       // node phase: [N][dx]
  float* dmbar;           // node phase: [N][dm]
  float* dxdst;           // edge phase: [N][dx], destination-side input gradient
  float* desrc;           // edge phase: [E][dx], per-edge source-side input gradient (CSR order)
  int dx;
  int need_dz0;
  int store_last;         // Z_L has to be recomputed (activation on the last layer, or max/min)
  int has_dst_side;
  int gout_ld;            // node phase, tensor-core kernels: leading dimension of gout_ptr (0: dout)
  const float* yact;      // node phase, tensor-core kernels: when set, the cotangent is multiplied by act'(y) on load, y = yact [N][dout]
  int yact_kind;          //   (an activation that follows the MLP's identity last layer: GCNConv's out >= in branch)
  int direct_src;         // tensor-core edge kernel, hoisted input: the source-side cotangent row of an edge IS its dZ_0 row (one SRC
                          // segment over all input rows): written from registers, no pass over the shared-memory tile
  int skip_w0;            // tensor-core kernels: layer 0 is the identity of a hoisted first layer -- its weight gradient is not formed
  int dst_c0, dst_w;      // likewise for the destination side: dxdst columns outside [dst_c0, dst_c0 + dst_w) are not written
  int src_c0, src_w;      // tensor-core edge kernels: x columns [src_c0, src_c0 + src_w) carry source-side cotangents; desrc is
                          // [E][src_w] (src_w == 0: all dx columns)
  int zoff[NGPDE_MAX_LAYERS + 1];
  int offG0, offG1, offW, offH, offDM, offP, offDZ, offDH, offRed;
  float* gno_S;        // contract == 2 (ngpde_gno.cuh): [N][gno_Ka * gin], rebuilt here for dB = S' DM
  const float* gno_T;  // contract == 2: [N][gno_Ka * gin] = DM B'
  int gno_Ka;
  int part_stride;     // floats per CTA in dparams_partial
  int offZt, offTs;
  int debug_skip;      // developer aid (NGPDE_OPT_DEBUG_SKIP): bitmask of phases left out, for phase timing only
};

// sign with which a segment's input gradient flows to the destination / source row
__device__ __forceinline__ float coef_dst(int kind) {
  return kind == SEG_DST ? 1.f : (kind == SEG_SMD ? -1.f : (kind == SEG_DMS ? 1.f : 0.f));
}
__device__ __forceinline__ float coef_src(int kind) {
  return kind == SEG_SRC ? 1.f : (kind == SEG_SMD ? 1.f : (kind == SEG_DMS ? -1.f : 0.f));
}

// bare Dense chains over the node axis (ngpde_conv.cu), used by GCNConv
int make_mlp_dev(const ngpde_mlp& m, MlpDev* out, const char* what);
size_t node_mlp_forward_ws(const MlpDev& mlp);
int node_mlp_forward(const ngpde_graph* g, const MlpDev& mlp, const float* params, const float* x, float* y,
                     cudaStream_t st, void* workspace = nullptr, size_t ws_bytes = 0, const float* snode = nullptr, int ds = 0,
                     int out_ld = 0);
size_t node_mlp_backward_ws(const ngpde_graph* g, const MlpDev& mlp);
int node_mlp_backward(const ngpde_graph* g, const MlpDev& mlp, const float* params, const float* x, const float* dy,
                      float* dx, float* dparams, void* workspace, size_t ws_bytes, cudaStream_t st, const float* snode = nullptr,
                      int ds = 0, int dy_ld = 0, const float* yact = nullptr, int yact_kind = 0);
// true when node_mlp_backward runs on the tensor-core kernel, which can apply act'(y) to the cotangent on load (`yact`)
bool node_mlp_backward_fuses_act(const ngpde_graph* g, const MlpDev& mlp);

}  // namespace ngpde
