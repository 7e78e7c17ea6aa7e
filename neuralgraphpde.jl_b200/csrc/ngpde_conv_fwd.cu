// Instantiations of the FFMA engine's forward kernels (edge and node phase, three tile sizes).
#include "ngpde_conv_kernels.cuh"
#include "ngpde_conv_launch.cuh"

namespace ngpde {
namespace {
template <bool NODE>
int launch_fwd(int te, const FwdArgs& a, int smem_bytes, int num_sms, cudaStream_t st) {
  int grid = 0;
  if (a.tg.n_units <= 0) return NGPDE_OK;
#define NGPDE_LAUNCH_FWD(TE)                                                          \
  {                                                                                   \
    if (int rc = launch_cfg(mp_fwd_kernel<TE, NODE>, smem_bytes, a.tg.n_units, num_sms, &grid)) return rc; \
    mp_fwd_kernel<TE, NODE><<<grid, NT, smem_bytes, st>>>(a);                         \
  }
  if (te == 128) NGPDE_LAUNCH_FWD(128) else if (te == 64) NGPDE_LAUNCH_FWD(64) else NGPDE_LAUNCH_FWD(32)
#undef NGPDE_LAUNCH_FWD
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}
}  // namespace

int launch_fwd_edge(int te, const FwdArgs& a, int smem_bytes, int num_sms, cudaStream_t st) {
  return launch_fwd<false>(te, a, smem_bytes, num_sms, st);
}
int launch_fwd_node(int te, const FwdArgs& a, int smem_bytes, int num_sms, cudaStream_t st) {
  return launch_fwd<true>(te, a, smem_bytes, num_sms, st);
}

}  // namespace ngpde
