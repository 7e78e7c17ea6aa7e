// Launchers of the FP32-FFMA engine's fused kernels.  The kernel templates (ngpde_conv_kernels.cuh) are instantiated in
// three translation units of their own -- ngpde_conv_fwd.cu, ngpde_conv_bwd_edge.cu, ngpde_conv_bwd_node.cu -- so that they
// compile in parallel and the host logic in ngpde_conv.cu rebuilds in seconds.
#pragma once
#include "ngpde_conv.cuh"

namespace ngpde {

int launch_fwd_edge(int te, const FwdArgs& a, int smem_bytes, int num_sms, cudaStream_t st);
int launch_fwd_node(int te, const FwdArgs& a, int smem_bytes, int num_sms, cudaStream_t st);
int bwd_grid_edge(int te, int smem_bytes, int n_units, int num_sms, int* grid);
int bwd_grid_node(int te, int smem_bytes, int n_units, int num_sms, int* grid);
int launch_bwd_edge(int te, const BwdArgs& a, int smem_bytes, int grid, cudaStream_t st);
int launch_bwd_node(int te, const BwdArgs& a, int smem_bytes, int grid, cudaStream_t st);

// shared by the launchers: opt the kernel into `smem_bytes` of dynamic shared memory and size a persistent grid
template <class K>
int launch_cfg(K kernel, int smem_bytes, int n_units, int num_sms, int* grid) {
  NGPDE_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  int occ = 0;
  NGPDE_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, NT, smem_bytes));
  if (occ < 1) {
    set_error("kernel cannot be resident with %d bytes of shared memory", smem_bytes);
    return NGPDE_ERR_UNSUPPORTED;
  }
  *grid = n_units < occ * num_sms ? (n_units > 1 ? n_units : 1) : occ * num_sms;
  return NGPDE_OK;
}

}  // namespace ngpde
