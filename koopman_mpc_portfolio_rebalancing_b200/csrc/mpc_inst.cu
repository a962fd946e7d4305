// One (H, APT) instantiation of the MPC / backtest kernels per translation unit, so that the variants
// compile in parallel (each takes ~1 min of ptxas time).  Built with -DKMPC_H=<H> -DKMPC_APT=<APT>.
#include "mpc_kernels.cuh"

#ifndef KMPC_H
#error "compile with -DKMPC_H=<horizon> -DKMPC_APT=<assets per lane> -DKMPC_NS=<slot stride>"
#endif
#define KMPC_CAT2(a, b, c, d) a##b##_##c##_##d
#define KMPC_CAT(a, b, c) KMPC_CAT2(a, _inst, b, c)
#define KMPC_APT ((KMPC_NS + 31) / 32)

namespace kmpc {
int KMPC_CAT(launch_mpc, KMPC_H, KMPC_NS)(const MpcSolveArgs& A, int sm_count, cudaStream_t st) {
  return launch_mpc<KMPC_H, KMPC_APT, KMPC_NS>(A, sm_count, st);
}
int KMPC_CAT(launch_bt, KMPC_H, KMPC_NS)(const BacktestArgs& A, int sm_count, cudaStream_t st) {
  return launch_bt<KMPC_H, KMPC_APT, KMPC_NS>(A, sm_count, st);
}
}  // namespace kmpc
