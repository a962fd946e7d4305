// Chooses the GEMM kernel for one launch: the tcgen05 3xTF32 kernel for the bulk shapes it supports, the fp32
// SIMT kernel otherwise.  Both are sm_100a CUDA kernels of this library; there is no other backend.
#include "gemm.cuh"
int kmpc_fail_cuda(cudaError_t e, const char* what);
namespace kmpc {
int launch_gemm(const GemmArgs& g, cudaStream_t st, long long* launches) {
  int rc = launch_gemm_tc(g, st);
  if (rc == -100) rc = launch_gemm_simt(g, st);
  if (launches) ++*launches;
  if (rc) return kmpc_fail_cuda((cudaError_t)rc, "gemm kernel");
  return 0;
}
}  // namespace kmpc
