// Instantiations of the FFMA engine's backward kernels, edge phase (three tile sizes).
#include "ngpde_conv_kernels.cuh"
#include "ngpde_conv_launch.cuh"

namespace ngpde {

int bwd_grid_edge(int te, int smem_bytes, int n_units, int num_sms, int* grid) {
  if (te == 128) return launch_cfg(mp_bwd_kernel<128, false>, smem_bytes, n_units, num_sms, grid);
  if (te == 64) return launch_cfg(mp_bwd_kernel<64, false>, smem_bytes, n_units, num_sms, grid);
  return launch_cfg(mp_bwd_kernel<32, false>, smem_bytes, n_units, num_sms, grid);
}

int launch_bwd_edge(int te, const BwdArgs& a, int smem_bytes, int grid, cudaStream_t st) {
  if (a.tg.n_units <= 0) return NGPDE_OK;
  if (te == 128) mp_bwd_kernel<128, false><<<grid, NT, smem_bytes, st>>>(a);
  else if (te == 64) mp_bwd_kernel<64, false><<<grid, NT, smem_bytes, st>>>(a);
  else mp_bwd_kernel<32, false><<<grid, NT, smem_bytes, st>>>(a);
  NGPDE_CUDA_TRY(cudaGetLastError());
  return NGPDE_OK;
}

}  // namespace ngpde
