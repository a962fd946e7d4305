// Definitions shared by the MPC kernels: the program, solver options, status codes, acceptance constants and warp
// helpers.  The program is the one of /root/reference/mpc.py:27-117 (solve_mpc_log_utility):
//
//   max  sum_k log(w_k . R_k) - lam * sum_k ||w_k - w_{k-1}||_1          (w_0 = current weights)
//   s.t. 1'w_k = 1,  w_k >= 0 (unless allow_short),  ||w_k - w_{k-1}||_1 <= tau (when tau > 0)
//
// Epigraph form with u_k >= |w_k - w_{k-1}|; slacks sp = u - d, sq = u + d, sc_k = tau - sum_i u_k.
// Mehrotra predictor-corrector.  The Newton system is solved through the problem structure:
//   * per asset, eliminating u leaves an SPD tridiagonal H x H system = a path network
//     (ground -e_1- w_1 -e_2- w_2 ...) whose Green's functions are built from series/parallel
//     conductances and multiplicative decay factors only (no cancellation, accurate when the
//     barrier weights span 1e-12 .. 1e+12);
//   * the couplings across assets (budget row, log-curvature R R', turnover cap) form a <= 3H border
//     whose SPD Schur complement K is assembled over the assets and factorised by one warp.
// The numpy twin of the solver (mpc_lane.cuh), iteration for iteration, is oracle/mpc_oracle.py::solve_structured.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <math_constants.h>
#include <stdint.h>

namespace kmpc {

enum : int { ST_OPTIMAL = 0, ST_INACCURATE = 1, ST_FAILED = 2, ST_NONFINITE = 3 };

struct IpmOptions {
  double tol;        // pres / gap tolerance
  double tol_dual;   // dual residual tolerance
  double delta;      // primal proximal regularisation of the Newton matrix
  double step_frac;  // fraction to the boundary
  double mu0;
  double dual_init;
  int max_iter;
  int second_attempt;     // 1: a solve that does not end "optimal" is repeated with the robust parameters (kRobust*)
  int clip_first_trade;   // 1: pull the executed trade of an optimal_inaccurate plan back onto the turnover cap
  int active_set;         // 1: backtests switch to reduced (active-set) solves once the portfolio has concentrated (host-side dispatch)
  int active_seg;         // decisions per work item of the active-set kernel (0: whole backtests)
};

__host__ __device__ inline IpmOptions default_ipm_options() {
  IpmOptions o;
  // step_frac / dual_init tuned on the 1.0 M-decision config-2 replay and a random instance mix (N 2..64, H 1..5,
  // lam 0..0.1, tau 0..1): (0.995, 3e-3) -> (0.9999, 1e-3) takes 9.35 -> 8.00 iterations per decision at the same
  // failure rate (3 fallbacks per million), 9.8 -> 8.9 on the mix with zero failures
  o.tol = 1e-10; o.tol_dual = 1e-8; o.delta = 1e-5; o.step_frac = 0.9999; o.mu0 = 1e-3; o.dual_init = 1e-3;
  o.max_iter = 100;
  o.second_attempt = 1;
  o.clip_first_trade = 1;
  o.active_set = 1;
  o.active_seg = 32;
  return o;
}

// Acceptance of an iterate the iteration could not push to the tolerances (iteration cap, breakdown of the border
// factorisation once the barrier weights span > 20 decades): "optimal_inaccurate" (mpc.py:113 uses such weights)
// when it is primal feasible, the gap has collapsed and the dual residual is at the level a first-order reference
// solver stops at (SCS eps 1e-4).  Measured on the config-2 replay: every such iterate is within 5e-7 relative of
// the optimal objective and 4e-4 of the optimal first-stage weights, whereas holding the weights (the fallback) is
// 0.1 away.
constexpr double kLoosePres = 1e-8, kLooseDres = 1e-4, kLooseGap = 1e-7;
// Second attempt of a solve whose first (aggressive) attempt did not end "optimal": restart from the cold starting point
// with textbook-robust parameters — one common primal/dual step length, a shorter fraction to the boundary, a centring
// floor and a weaker proximal term (oracle/mpc_oracle.py ROBUST_*; the parameterisation of the dense oracle).  On the
// ~50 decisions per million of a config-2 step that the first attempt leaves "optimal_inaccurate" (near-degenerate
// optima: the dual residual stalls at 1e-7..1e-5 while the gap collapses; a few of them 2-5e-6 off the optimal
// objective) the second attempt reaches "optimal" on every one.
constexpr double kRobustStepFrac = 0.995, kRobustSigmaMin = 0.05, kRobustDelta = 1e-7;
constexpr int ST_RESTART = -2;          // LaneIpm::check(): call begin() again for the second attempt
// Mehrotra's second-order term is scaled by min(1, affine step / kCorrFull): see oracle/mpc_oracle.py (CORRECTOR_FULL_STEP)
constexpr double kCorrFull = 0.3;
// lane kernel: factorisation breakdowns answered by a stronger proximal term before the decision falls back
constexpr int kMaxFactorRetries = 4;

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(kFull, v, m); }
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(kFull, v, src); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
  return v;
}

}  // namespace kmpc
