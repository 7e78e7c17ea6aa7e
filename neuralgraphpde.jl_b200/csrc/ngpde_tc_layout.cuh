// Shapes, TMEM / shared-memory maps and host entry points of the tensor-core (tcgen05) message-passing kernels.
// The kernels themselves live in ngpde_tc.cuh / ngpde_tc_bwd.cuh and are compiled in their own translation unit
// (ngpde_tc.cu); ngpde_conv.cu only sees this header.
#pragma once
#include "ngpde_conv.cuh"

namespace ngpde {

constexpr int TC_TILE = 128;    // rows per tile = MMA M
constexpr int TC_GTHREADS = 256;  // threads per group: 8 warps = 4 TMEM lane quarters x 2 column halves
constexpr int TC_GROUPS = 2;    // groups per CTA, each with its own tile in flight
constexpr int TC_MAXN = 64;     // widest layer output the path accepts
constexpr int TC_THREADS = TC_GTHREADS * TC_GROUPS;

// Shapes of one MLP on the tensor-core path.  Layer l reads A columns [0, Kd) (data, zero padded), a "ones" block at
// column Kd (1, 0, ..., 0) that multiplies the bias row of the weight image -- so the bias add costs no epilogue
// instruction -- and writes Np = pad16(N) accumulator columns.
struct TcLayout {
  int L;
  int K[NGPDE_MAX_LAYERS], N[NGPDE_MAX_LAYERS];    // logical layer shapes
  int Kd[NGPDE_MAX_LAYERS];                        // pad16(K): data columns / rows
  int Kp[NGPDE_MAX_LAYERS];                        // Kd + 8: rows of the weight image (K extent of the MMAs)
  int Np[NGPDE_MAX_LAYERS];                        // pad16(N)
  int img_off[NGPDE_MAX_LAYERS];                   // float offset of the hi image; the lo image follows it
  int img_floats[NGPDE_MAX_LAYERS];                // floats of one image = ceil(Np/32) * Kp * 32
  int block_floats;                                // all images
  int kmax;                                        // widest Kp
  int cols_group;                                  // TMEM columns per group: TC_MAXN (D) + 2*kmax (A hi, A lo)
  int tmem_cols;                                   // allocation: power of two >= 32
  // Node phase: the input segments of layer 0 may be re-ordered on chip (16-column-aligned segments first, so that a
  // thread's 16-column chunk is 4 aligned float4 of ONE array).  Position p of the on-chip order holds original input row
  // po[i] + (p - ps[i]) for the segment i with ps[i] <= p < ps[i] + pw[i]; np == 0: identity.
  int np;
  int ps[8], po[8], pw[8];
};
constexpr int TC_MAX_CHUNKS = 16;  // 16-column chunks of a layer-0 input (Kd0 <= 256)

__host__ __device__ inline int tc_orig_row(const TcLayout& lay, int p) {
  for (int i = 0; i < lay.np; ++i)
    if (p >= lay.ps[i] && p < lay.ps[i] + lay.pw[i]) return lay.po[i] + (p - lay.ps[i]);
  return p;
}

constexpr int TCB_WORKERS = 512;  // 16 worker warps: thread = (tile row, 16-column chunk)
constexpr int TCB_THREADS = TCB_WORKERS + 32;  // + one dedicated MMA-issuing warp
constexpr int TCB_MAXL = 4;     // register accumulators for at most 4 layers
constexpr int TCB_HALF = 64;    // rows per weight-gradient staging pass

// Backward on tensor cores: TMEM column map, shared-memory map and eligibility (see ngpde_tc_bwd.cuh).
struct TcBwdPhase {
  bool on = false;
  TcLayout lay{};
  int smem = 0;
  int c_zs[NGPDE_MAX_LAYERS] = {0};
  int c_a = 0, a_width = 0, c_d = 0, c_dw = 0, c_d0 = 0, c_dw0 = 0, tmem_cols = 0, dw_alt = 0;
  int off_cols = 0, off_stage = 0, off_dz = 0, nzh = 0, nzl = 0;
  // shared-memory placement of the weight images: float offset of layer l's hi image; full = whole-tile staging of the
  // weight-gradient operands, which may need layers stream_a / stream_b to share one slot (ngpde_tc_bwd.cuh)
  int woff[NGPDE_MAX_LAYERS] = {0};
  int stream_a = -1, stream_b = -1;
  bool full = false;
  size_t ws_off = 0;
  int grid = 0;
};

struct TcPhase {
  bool on = false;
  TcLayout lay{};
  int smem = 0, off_cols = 0, off_groups = 0, group_bytes = 0;
  size_t ws_off = 0;  // byte offset of the prepared weight block in the forward workspace
};


// eligibility + layout (false: the phase runs on the FP32-FFMA engine instead)
bool tc_bwd_make(const MlpDev& m, bool contract, bool addend, bool node, int aggr, bool need_dz0, TcBwdPhase* t);
bool tc_make_layout(const MlpDev& m, bool contract, bool addend, int dout, bool node, TcLayout* lay, int* smem_bytes,
                    int* off_cols, int* off_groups, int* group_bytes);
// launches (weight-image preparation + the fused kernel) on `st`
int launch_fwd_tc(bool node, int num_sms, const TcPhase& t, const MlpDev& mlp, const float* params, const FwdArgs& base,
                  float* wblock, cudaStream_t st);
int launch_bwd_tc(bool node, const TcBwdPhase& t, const MlpDev& mlp, const BwdArgs& base, float* wblock, cudaStream_t st);
void tc_set_enabled(bool on);          // NGPDE_OPT_TENSOR_CORES
long long* tc_get_debug_buffer();
void tc_set_debug_buffer(long long* p);  // phase timestamps of the edge-phase backward (tools/tcb_phases.py)

}  // namespace ngpde
