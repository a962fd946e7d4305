// One (H, G) instantiation of the lane-per-asset MPC / backtest kernels per translation unit.
// Built with -DKMPC_H=<H> -DKMPC_G=<warps per problem>.
#include "mpc_lane_kernels.cuh"
#ifndef KMPC_H
#error "compile with -DKMPC_H=<horizon> -DKMPC_G=<warps per problem>"
#endif
#define KMPC_CAT2(a, b, c, d) a##b##_##c##_##d
#define KMPC_CAT(a, b, c) KMPC_CAT2(a, _lane, b, c)
namespace kmpc {
int KMPC_CAT(launch_mpc, KMPC_H, KMPC_G)(const MpcSolveArgs& A, int sm_count, cudaStream_t st) {
  return launch_mpc_lane<KMPC_H, KMPC_G>(A, sm_count, st);
}
int KMPC_CAT(launch_bt, KMPC_H, KMPC_G)(const BacktestArgs& A, int sm_count, cudaStream_t st) {
  return launch_bt_lane<KMPC_H, KMPC_G>(A, sm_count, st);
}
#if KMPC_G == 1 && (KMPC_H <= 5 || KMPC_H == 10)
// the active-set kernel (one warp per reduced problem) lives with the one-warp variant of its horizon
int KMPC_CAT(launch_bta, KMPC_H, KMPC_G)(const BacktestArgs& A, int sm_count, cudaStream_t st) {
  return launch_bt_active<KMPC_H>(A, sm_count, st);
}
#endif
#if KMPC_G == 4 && (KMPC_H == 5 || KMPC_H == 10)
// the wide active-set kernel (KMPC_WIDE_G warps per reduced problem, universes beyond 128 assets) lives with the four-warp variant
int KMPC_CAT(launch_btaw, KMPC_H, KMPC_G)(const BacktestArgs& A, int sm_count, cudaStream_t st) {
  return launch_bt_active_wide<KMPC_H>(A, sm_count, st);
}
int KMPC_CAT(wide_threads, KMPC_H, KMPC_G)() { return 32 * kWideG; }      // threads = active assets at most of a wide problem
#endif
}  // namespace kmpc
