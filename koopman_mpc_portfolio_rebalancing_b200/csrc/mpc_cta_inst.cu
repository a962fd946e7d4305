// One (H, G) instantiation of the CTA-per-problem MPC / backtest kernels per translation unit.
// Built with -DKMPC_H=<H> -DKMPC_G=<asset groups of 32 per stage>.
#include "mpc_cta_kernels.cuh"
#ifndef KMPC_H
#error "compile with -DKMPC_H=<horizon> -DKMPC_G=<asset groups>"
#endif
#define KMPC_CAT2(a, b, c, d) a##b##_##c##_##d
#define KMPC_CAT(a, b, c) KMPC_CAT2(a, _cta, b, c)
namespace kmpc {
int KMPC_CAT(launch_mpc, KMPC_H, KMPC_G)(const MpcSolveArgs& A, int sm_count, cudaStream_t st) {
  return launch_mpc_cta<KMPC_H, KMPC_G>(A, sm_count, st);
}
int KMPC_CAT(launch_bt, KMPC_H, KMPC_G)(const BacktestArgs& A, int sm_count, cudaStream_t st) {
  return launch_bt_cta<KMPC_H, KMPC_G>(A, sm_count, st);
}
}  // namespace kmpc
