// Kernels around the lane-per-asset IPM solver (mpc_lane.cuh): mpc_solve = mpc.py:27-117, backtest =
// backtest.py:173-249.  mpc_solve: one block of G warps per
// problem; backtest: a persistent block hosts several backtests ("slots" of G warps each) that walk through the
// Newton iteration together.  Thread i of a slot = asset i, all stages of an asset in that thread's registers.
// backtest_lane_kernel solves all N assets at every decision; backtest_active_kernel (further down) solves the held
// assets + candidates on one warp and verifies the rest, and takes over once a portfolio has concentrated.
#pragma once
#include "kmpc_internal.cuh"
#include "mpc_lane.cuh"
#include <stdio.h>
#include <type_traits>

// register budget of the backtest kernel: __maxnreg__ and __launch_bounds__ are mutually exclusive
#ifdef KMPC_LANE_MAXNREG
#define KMPC_LANE_BT_ATTR(threads) __maxnreg__(KMPC_LANE_MAXNREG)
#else
#define KMPC_LANE_BT_ATTR(threads) __launch_bounds__(threads, 1)
#endif
#ifndef KMPC_LANE_SYNC_EVERY
#define KMPC_LANE_SYNC_EVERY 4    // block barrier every 4th trip (measured per config-2 step, final code: 3: 195.3 ms, 4: 193.0, 6: 196.5; earlier 2: 204, 4: 196, 8: 201; before the start-up code shrank 1: 215, 4: 208, 16: 215, 64: 231)
#endif
#ifndef KMPC_LANE_MINB
#define KMPC_LANE_MINB 1      // resident blocks per SM the register allocation is sized for (0/1 = no cap)
#endif

namespace kmpc {

__device__ __forceinline__ float exp_cr32_lane(float y) { return __double2float_rn(exp((double)y)); }

// FIX kernels (structure flags compile-time, see LaneIpm) and generic kernels are launched as a pair when per-problem
// lam / tau arrays are given: `want` selects which of the two does the work once lane_flags_kernel has looked at
// the arrays (no host synchronisation); the other one exits at once.
static __global__ void lane_flags_kernel(const double* lam, const double* tau, double lam0, double tau0, int n, int* flag) {
  __shared__ int bad;
  if (threadIdx.x == 0) bad = 0;
  __syncthreads();
  int b = 0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double l = lam ? lam[i] : lam0, t = tau ? tau[i] : tau0;
    if (!(l > 0.0) || !(t > 0.0)) b = 1;
  }
  if (b) bad = 1;
  __syncthreads();
  if (threadIdx.x == 0) *flag = bad ? 0 : 1;
}

template <int H, int G, bool FIX>
__global__ void __launch_bounds__(32 * G, KMPC_LANE_MINB)
mpc_solve_lane_kernel(MpcSolveArgs A, int want) {
  using Ipm = LaneIpm<H, G, (G > 4 || H > 5), FIX>;
  extern __shared__ double smem[];
  if (want >= 0 && *A.fix_flag != want) return;
  Ipm s;
  s.bind(smem, A.N, 0);
  const int N = A.N;
  const IpmOptions opt = A.opt;
  for (int p = blockIdx.x; p < A.P; p += gridDim.x) {
    double w0 = 0.0;
    if (s.valid) {
      w0 = A.w_cur[(size_t)p * N + s.tid];
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const size_t idx = ((size_t)p * H + k) * N + s.tid;
        s.R[k] = A.yhat ? (double)exp_cr32_lane(A.yhat[idx]) : exp(A.yhat64[idx]);
      }
    }
    const double lam = A.lam ? A.lam[p] : A.lam0;
    const double tau = A.tau ? A.tau[p] : A.tau0;
    int iters; double kkt[3];
    const int st = s.solve(w0, N, lam, tau, A.allow_short != 0, opt, iters, kkt);
    double val = CUDART_NAN;
    if (st <= ST_INACCURATE) { if (opt.clip_first_trade) s.clip_first_trade(w0); val = s.objective(w0); }
    if (s.valid) {
#pragma unroll
      for (int k = 0; k < H; ++k) A.w_out[((size_t)p * H + k) * N + s.tid] = s.w[k];
    }
    if (threadIdx.x == 0) {
      if (A.obj) A.obj[p] = val;
      if (A.kkt) { A.kkt[3 * p] = kkt[0]; A.kkt[3 * p + 1] = kkt[1]; A.kkt[3 * p + 2] = kkt[2]; }
      if (A.status) A.status[p] = st;
      if (A.iters) A.iters[p] = iters;
    }
    s.sync();
  }
}

// Persistent backtest kernel.  A block hosts P independent backtests ("slots", G warps each).  The solver code is
// ~8 k straight-line instructions per Newton iteration, several times the 32 KB instruction cache of an SM: when
// every resident problem walks through it at its own pace the warps starve on instruction fetch (ncu: 41 % of
// all stall samples `no_instruction`, throughput 2.15x going from 1 to 4 independent blocks per SM).  Here all
// slots of the SM pass through the phases of an iteration TOGETHER (a block-wide barrier every few trips keeps them
// within a phase or two of each other), so one fetched line feeds every warp; a slot whose decision has converged books the portfolio step and starts its next
// decision inside the same trip, so no slot ever idles through a phase.
// Book-keeping of one slot's current backtest (backtest.py:161-217, 221-249).  Lives in shared memory and is
// touched by thread 0 of the slot only: as registers it would cost every thread of the kernel ~30 registers.
struct SlotBook {
  double V, ccoef, mean, m2, cum, peak, maxdd, sum_turn, v_first;
  long long it_total;
  int n, n_opt, n_inacc, n_fail;
};
// A backtest that changes kernels (active-set pipeline) travels as state[b] = weights [N] | book-keeping [14] | step index.
__device__ __forceinline__ void book_save(const SlotBook& k, int t, double* S) {
  S[0] = k.V; S[1] = k.ccoef; S[2] = k.mean; S[3] = k.m2; S[4] = k.cum; S[5] = k.peak; S[6] = k.maxdd; S[7] = k.sum_turn;
  S[8] = k.v_first; S[9] = (double)k.it_total; S[10] = k.n; S[11] = k.n_opt; S[12] = k.n_inacc; S[13] = k.n_fail; S[14] = t;
}
__device__ __forceinline__ void book_load(SlotBook& k, const double* S) {
  k.V = S[0]; k.ccoef = S[1]; k.mean = S[2]; k.m2 = S[3]; k.cum = S[4]; k.peak = S[5]; k.maxdd = S[6]; k.sum_turn = S[7];
  k.v_first = S[8]; k.it_total = (long long)S[9]; k.n = (int)S[10]; k.n_opt = (int)S[11]; k.n_inacc = (int)S[12]; k.n_fail = (int)S[13];
}
// a weight at or below this counts as "not held" (an interior-point solve leaves ~1e-11 on assets at their bound)
constexpr double kHeldThr = 1e-9;
// One decision of backtest b goes into the books (thread 0 of the slot): trading cost, portfolio value, the history row and
// the running statistics of backtest.py:175-217; at the last step calculate_metrics (backtest.py:221-249).
__device__ __forceinline__ void book_decision(SlotBook& k, const BacktestArgs& A, int b, int st, long long iters, double turnover,
                                              double port_ret, bool market, bool last) {
  k.it_total += iters;
  k.n_opt += (st == ST_OPTIMAL); k.n_inacc += (st == ST_INACCURATE); k.n_fail += (st >= ST_FAILED);
  const double cost = k.ccoef * turnover * k.V;
  double V = k.V - cost;
  if (market) V *= (1.0 + port_ret);
  k.V = V;
  if (A.history) {
    double* hrow = A.history + ((size_t)b * A.n_hist + k.n) * 4;
    hrow[0] = V; hrow[1] = port_ret; hrow[2] = turnover; hrow[3] = cost;
  }
  if (k.n == 0) k.v_first = V;
  const int n = ++k.n;
  const double dlt = port_ret - k.mean;
  k.mean += div_fast(dlt, (double)n);
  k.m2 += dlt * (port_ret - k.mean);
  k.cum *= (1.0 + port_ret);
  k.peak = fmax(k.peak, k.cum);
  k.maxdd = fmin(k.maxdd, div_fast(k.cum - k.peak, k.peak));
  k.sum_turn += turnover;
  if (__builtin_expect(last, 0)) {                     // calculate_metrics (backtest.py:221-249)
    double* m = A.metrics + (size_t)b * 5;
    const double inv_n = rcp_fast((double)n);
    const double sd = sqrt(k.m2 * inv_n);
    m[0] = div_fast(sqrt(252.0) * k.mean, sd + 1e-8);
    m[1] = k.maxdd;
    m[2] = k.sum_turn * inv_n;
    m[3] = V;
    m[4] = div_fast(V, k.v_first) - 1.0;
    if (A.solve_stats) {
      long long* ss = A.solve_stats + (size_t)b * 4;
      ss[0] = k.n_opt; ss[1] = k.n_inacc; ss[2] = k.n_fail; ss[3] = k.it_total;
    }
  }
}

// The portfolio step of one decision for my asset (backtest.py:175-217): executes the first trade of the plan held in s.w[0]
// from the weight wc, returns the drifted weight (the new wc) and, in every thread, the turnover and the portfolio's return
// of the day.  e_next = exp(realised log-return of my asset on the next day), float32 as the reference computes it.
template <typename Ipm>
__device__ __forceinline__ double portfolio_step(Ipm& s, const IpmOptions& opt, double wc, float e_next, bool market, double& turnover,
                                                 double& port_ret) {
  double wn = s.valid ? s.w[0] : 0.0;                                                // backtest.py:131
  float r32 = 0.0f;
  if (s.valid && market) r32 = __fsub_rn(e_next, 1.0f);                              // backtest.py:193
  double v[3] = {fabs(wn - wc), wn * (double)r32, wc * (double)r32}, T[3];
  s.sync();
  s.template block_sum<3>(v, T);
  turnover = T[0];
  port_ret = market ? T[1] : 0.0;
  if (opt.clip_first_trade && s.tau > 0.0 && turnover > s.tau) {
    // An iterate accepted after the factorisation broke down next to the optimum (status optimal_inaccurate)
    // can sit ~1e-5 outside the turnover cap, whose slack the iteration does not re-derive from w: pull the
    // trade back onto the cap along its own direction (budget and sign constraints are kept).
    const double sc = div_fast(s.tau, turnover);
    wn = fma(sc, wn - wc, wc);
    if (market) port_ret = fma(sc, T[1] - T[2], T[2]);
    turnover = s.tau;
  }
  if (!market) return wn;
  double denom = 1.0 + port_ret;
  if (fabs(denom) < 1e-8) denom = 1e-8;
  return div_fast(wn * (double)__fadd_rn(1.0f, r32), denom);                         // (1.0 + f32) stays f32
}

template <int H, int G, int P, bool FIX>
__global__ void KMPC_LANE_BT_ATTR(32 * G * P)
backtest_lane_kernel(BacktestArgs A, int want) {
  using Ipm = LaneIpm<H, G, (G > 4 || H > 5), FIX>;
  extern __shared__ double smem[];
  if (want >= 0 && *A.fix_flag != want) return;
  __shared__ int next_b[P];
  __shared__ SlotBook books[P];
  const int slot = __shfl_sync(kFull, (int)threadIdx.x / (32 * G), 0);
  Ipm s;
  s.bind(smem + (size_t)slot * Ipm::SMEM_DOUBLES, A.N, slot);
  const int N = A.N;
  const IpmOptions& opt = A.opt;
  // ---- per-thread state of the slot's current backtest ----------------------------------------------------------
  int b = 0, t = 0, t_in = -1;                     // t_in: step at which a resumed backtest came in (it runs at least one decision here)
  double wc = 0.0;                                 // my asset's current weight
  float e_next = 1.0f;                             // exp(realised log-return of my asset on the day after the decision)
  auto fetch = [&]() -> bool {                     // next backtest of this slot (dynamic: iteration counts differ)
    for (;;) {
      if (s.tid == 0) next_b[slot] = atomicAdd(A.work_counter, 1);
      s.sync();
      b = __shfl_sync(kFull, next_b[slot], 0);
      if (b >= A.B) return false;
      if (A.phase < 2 || A.bt_status[b] == 2) break;                                   // resume passes: suspended backtests only
      s.sync();                                                                        // everybody has read next_b
    }
    if (A.phase >= 2) {                                                                // carry on where the active-set kernel stopped
      const double* S = A.state + (size_t)b * A.state_ld;
      wc = s.valid ? S[s.tid] : 0.0;
      t = (int)S[N + 14];
      t_in = t;
      if (s.tid == 0) book_load(books[slot], S + N);
      return true;
    }
    t_in = -1;
    wc = s.valid ? 1.0 / (double)N : 0.0;                                              // backtest.py:161
    t = 0;
    if (s.tid == 0) {
      SlotBook& k = books[slot];
      k.V = A.capital ? A.capital[b] : A.capital0;
      k.ccoef = A.cost_coeff ? A.cost_coeff[b] : A.cost_coeff0;
      k.mean = 0.0; k.m2 = 0.0; k.cum = 1.0; k.peak = -CUDART_INF; k.maxdd = CUDART_INF; k.sum_turn = 0.0; k.v_first = 0.0;
      k.it_total = 0; k.n = 0; k.n_opt = 0; k.n_inacc = 0; k.n_fail = 0;
    }
    return true;
  };
  bool active = (A.n_steps > 0) ? fetch() : false;
  bool need_start = true;
  int st = -1;
#ifdef KMPC_LANE_PROFILE
  __shared__ long long prof[12];
  if (threadIdx.x == 0) { for (int i = 0; i < 12; ++i) prof[i] = 0; s.prof_ = prof; s.tl_ = clock64(); }
#endif
  __syncthreads();
#pragma unroll 1
  for (unsigned trip = 0;; ++trip) {
    if (uni(active)) {
#pragma unroll 1
      for (;;) {
        if (uni(need_start) || uni(st == ST_RESTART)) {
          const bool restart = !need_start;                          // second attempt: same returns, same weights
          if (uni(!restart && (A.phase == 1 || A.phase == 3) && t > 0 && t != t_in)) {
            // dense start of the active-set pipeline: once few assets are held, the backtest moves to the kernel that
            // solves reduced problems (one warp per problem)
            s.sync();
            const double held = s.block_sum1((s.valid && wc > kHeldThr) ? 1.0 : 0.0);
            if (uni(held <= (double)A.as_hmax)) {
              double* S = A.state + (size_t)b * A.state_ld;
              if (s.valid) S[s.tid] = wc;
              if (s.tid == 0) {
                book_save(books[slot], t, S + N);
                A.bt_status[b] = 1;
                if (A.phase == 3) atomicSub(A.done_counter, 1);                        // it had been counted out when it was suspended
                A.ready_ring[(unsigned)atomicAdd(A.queue_ctr + 1, 1) % (unsigned)A.B] = b;      // ready queue of the active-set kernel
              }
              __syncwarp();
              s.sync();
              active = fetch();
              if (!active) break;
              continue;
            }
          }
          if (!restart) {
            const size_t yb = (size_t)(A.yhat_index ? A.yhat_index[b] : b) * A.yhat_stride;
            const size_t rb = (size_t)(A.realized_index ? A.realized_index[b] : b) * A.realized_stride;
            const float y_next = (s.valid && t + 1 < A.rows) ? A.realized[rb + (size_t)(t + 1) * N + s.tid] : 0.0f;
            e_next = s.load_returns(A.yhat + yb + (size_t)t * H * N, (size_t)N, y_next, s.tid);    // mpc.py:55
          } else {
            s.sync();
          }
          st = s.begin(wc, N, A.lam ? A.lam[b] : A.lam0, A.tau ? A.tau[b] : A.tau0, A.allow_short != 0, opt, restart);
          need_start = false;
        }
        if (uni(st == -1)) st = s.check(opt);
        if (uni(st == ST_RESTART)) continue;                       // the first attempt did not end "optimal"
        if (uni(st < 0)) break;                                    // take a Newton step
        // ---- the decision is made: portfolio step (backtest.py:175-217) ---------------------------------------
        const bool market = (t + 1 < A.rows);
        double turnover, port_ret;
        wc = portfolio_step(s, opt, wc, e_next, market, turnover, port_ret);
        t += A.rebalance_freq;
        const bool last = (t >= A.n_steps);
        if (s.tid == 0) {
          book_decision(books[slot], A, b, st, s.it_, turnover, port_ret, market, last);
        }
        need_start = true; st = -1;
        if (uni(last)) {
          if (A.final_weights && s.valid) A.final_weights[(size_t)b * N + s.tid] = wc;
          if (A.bt_status && s.tid == 0) { A.bt_status[b] = 3; if (A.phase == 1) atomicAdd(A.done_counter, 1); }
          active = fetch();
          if (!active) break;
        }
      }
    }
    // A block barrier every few trips keeps the slots within a phase or two of each other, which is what the
    // instruction cache needs; a barrier per trip makes every slot wait for the one that books a decision (19 % of
    // all stall samples).  Measured: barriers between the phases change nothing, staggering the slots half a trip
    // apart is slower (357 vs 273 ms).
    // (the barrier doubles as the exit vote: all slots out of work)
    KMPC_PROF(s, 0)
#ifdef KMPC_AS_DEBUG
    if (trip > 30000u) {
      if (s.lane == 0) printf("G kernel stuck: block %d slot %d warp %d phase %d active %d b %d t %d st %d need_start %d it %d\n",
                              (int)blockIdx.x, slot, s.warp, A.phase, (int)active, b, t, st, (int)need_start, s.it_);
      break;
    }
#endif
    if ((trip % KMPC_LANE_SYNC_EVERY) == 0) {
      if (__syncthreads_and(!active)) break;
    }
    KMPC_PROF(s, 1)
    const bool act_u = uni(active);                            // provably warp-uniform (see uni())
    bool ok = false;
    if (act_u) {
      s.factor_a();
      KMPC_PROF(s, 3)
      ok = s.factor_b();
      KMPC_PROF(s, 4)
    }
    if (ok) {
#pragma unroll 1
      for (int phase = 0; phase < 2; ++phase) {
        s.newton_phase(phase, opt);
        KMPC_PROF(s, 5 + phase)
      }
    }
  }
#ifdef KMPC_LANE_PROFILE
  if (threadIdx.x == 0 && blockIdx.x < 2)
    printf("lane_prof block %d: check/book/begin %lld  barrier %lld  factor_a.sweeps %lld  factor_a.K %lld  factor_b %lld  predictor.rest %lld  corrector.rest %lld  newton.sweep1+reduce %lld  newton.ksolve %lld\n",
           (int)blockIdx.x, prof[0], prof[1], prof[2], prof[3], prof[4], prof[5], prof[6], prof[7], prof[8]);
#endif
  // backtests without any step: NaN metrics (the host never asks for this; kept for completeness)
  if (A.n_steps <= 0) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < A.B * 5; q += gridDim.x * blockDim.x) A.metrics[q] = CUDART_NAN;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Active-set backtest kernel.  Once a portfolio has concentrated, the optimum of mpc.py:49-104 lives on a handful of assets:
// on the config-2 replay 2-7 of 50 assets are held after the first dozen decisions and the plan never touches more than 8.
// A slot here is ONE warp (LaneIpm<H, 1>: no problem-wide named barriers, eight problems per SM instead of four; measured
// 10.3 against 18.1 ns per Newton iteration and GPU) that solves the program RESTRICTED to an active set S of at most 32
// assets and then PROVES the restriction harmless:
//   S        = assets held (weight > kHeldThr; lighter ones are dropped, < 5e-8 of weight in total) + the two best forecasts
//              of every stage, compacted into the lanes of the warp;
//   solve    = the same interior-point iteration on |S| assets (fewer iterations as well: 6.9 against 8.2 on the replay);
//   verify   = an excluded asset i stays at w_ik = 0 for all stages iff there are y_k in [-c_k, c_k], c_k = lam + zc_k (the
//              subgradients of |w_ik - w_i,k-1| at 0 under the cost and the cap's dual) with
//                  nu_k - R_ik / rho_k + y_k - y_k+1 >= 0   for every stage k       (stationarity with zw_ik >= 0)
//              where nu, rho, zc are the duals / portfolio returns of the reduced solution.  Backwards from y_H+1 = 0 the
//              smallest admissible y_k = max(-c_k, y_k+1 - g_k) is also the best choice for stage k - 1, so the greedy
//              recursion decides feasibility exactly; almost every asset passes the one-comparison pre-check g_k > 0 for all
//              k, i.e. yhat_ik < log(nu_k rho_k).  The reduced plan padded with zeros then satisfies the KKT conditions of
//              the FULL convex program: it is an optimum of it, not an approximation;
//   repair   = an asset that fails joins S and the decision is solved again (never needed on the replay); more than 32
//              assets in S suspend the backtest: its state is saved and the full-width kernel finishes it (phase 2).
// Backtests arrive from phase 1 of backtest_lane_kernel (state[b]: weights, book-keeping, step index).
constexpr double kVerifyTol = 1e-8;     // on the subgradient bound; the duals of the reduced solution carry tol_dual = 1e-8
// H = 10 one-warp problems: sweep factors and corrector targets in shared memory (45 KB per slot, four slots per SM) rather
// than thread-private (LOC: 3 KB of local memory per thread, eight slots per SM)
#ifndef KMPC_ACTIVE_LOC
#define KMPC_ACTIVE_LOC(H) false
#endif
// Slots per block of the reduced-solve kernel.  G = 1: one warp per problem (eight problems per SM; H = 10 with the sweep
// factors in shared memory: four).  G > 1 ("wide": universes beyond 128 assets, where the full-width solver needs 16 warps
// per problem at 128 registers per thread): KMPC_WIDE_G warps per problem at the full register budget, 8 / G problems per SM,
// the solver in the layout of backtest_lane_kernel<H, G> (H = 10: thread-private sweep factors).
template <int H, int G = 1> struct ActiveSlots {
  static constexpr int P = (G > 1) ? (8 / G) : ((H > 5 && !KMPC_ACTIVE_LOC(H)) ? 4 : 8);
  static constexpr bool LOC = (G > 1) ? (H > 5) : KMPC_ACTIVE_LOC(H);
};
#ifndef KMPC_ACTIVE_SYNC_EVERY
#define KMPC_ACTIVE_SYNC_EVERY 4        // block barrier (and exit vote) every n-th trip of the active-set kernel
#endif

template <int H, int P, bool FIX, int NQ, int G = 1>
__global__ void __launch_bounds__(32 * G * P, 1)
backtest_active_kernel(BacktestArgs A, int want) {
  using Ipm = LaneIpm<H, G, ActiveSlots<H, G>::LOC, FIX>;
  constexpr int NT = 32 * G;                       // threads (= active assets at most) per problem
  constexpr int MAXQ = NQ;                         // assets per thread in passes over the whole universe (N <= NT NQ)
  extern __shared__ double smem[];
  if (want >= 0 && *A.fix_flag != want) return;
  __shared__ SlotBook books[P];
  __shared__ int xch[P][2][8];                     // G > 1: exchange between the warps of a slot (two alternating rows)
  const int slot = __shfl_sync(kFull, (int)threadIdx.x / NT, 0);
  const int lane = (int)threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  Ipm s;
  s.bind(smem + (size_t)slot * Ipm::SMEM_DOUBLES, NT, slot);
  const int tid = s.tid;                           // thread of the problem: [0, NT)
  double* wfull = smem + (size_t)P * Ipm::SMEM_DOUBLES + (size_t)slot * (NT * MAXQ);       // weights of all N assets
  int* sid = reinterpret_cast<int*>(smem + (size_t)P * (Ipm::SMEM_DOUBLES + NT * MAXQ)) + slot * NT;   // thread -> asset
  const int N = A.N;
  // the decision's forecasts [H][N] and the next day's realised returns [N], staged once per decision: the selection, the
  // solver's inputs and the verification all read them, and the global loads of a decision are in flight together
  // (large universes — config 3: 11 x 500 floats per slot — read them from global memory / L2 instead)
  constexpr bool staged = (G == 1) && NQ <= 4;     // N <= 128 (the launcher picks NQ from N)
  float* ystage = reinterpret_cast<float*>(reinterpret_cast<int*>(smem + (size_t)P * (Ipm::SMEM_DOUBLES + NT * MAXQ)) + P * NT) +
                  (size_t)slot * (H + 1) * N;
  const float* ysrc = ystage;                      // [H][N] forecasts of the decision (staged: fixed, in shared memory)
  const float* rnext = ystage + H * N;             // [N] next day's realised log-returns
  const IpmOptions& opt = A.opt;
  int b = 0, t = 0, count = 0, a = 0, extra_it = 0, seg_left = 0;
  unsigned member = 0;                             // bit q: asset tid + NT q is in S
  double wc = 0.0;
  float e_next = 1.0f;
  const float* yrow = nullptr;
  size_t rb = 0;
  double lam_b = 0.0, tau_b = 0.0;

  // ---- collectives over the threads of the slot (G = 1: the warp; G > 1: through xch and the slot's named barrier; every
  //      thread of the slot calls them in the same order, from slot-uniform control flow) -----------------------------------
  int xsel = 0;
  auto slot_sync = [&]() { s.sync(); };
  auto slot_bcast0 = [&](int v) -> int {           // the value of thread 0 of the slot
    if (G == 1) return __shfl_sync(kFull, v, 0);
    xsel ^= 1;
    if (tid == 0) xch[slot][xsel][0] = v;
    s.sync();
    return xch[slot][xsel][0];
  };
  auto slot_any = [&](bool c) -> bool {
    const bool w = __any_sync(kFull, c) != 0;
    if (G == 1) return w;
    xsel ^= 1;
    if (lane == 0) xch[slot][xsel][s.warp] = w ? 1 : 0;
    s.sync();
    int r = 0;
#pragma unroll
    for (int g = 0; g < G; ++g) r |= xch[slot][xsel][g];
    return r != 0;
  };
  auto slot_maxf = [&](float m) -> float {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    if (G == 1) return m;
    xsel ^= 1;
    if (lane == 0) xch[slot][xsel][s.warp] = __float_as_int(m);
    s.sync();
    float r = __int_as_float(xch[slot][xsel][0]);
#pragma unroll
    for (int g = 1; g < G; ++g) r = fmaxf(r, __int_as_float(xch[slot][xsel][g]));
    return r;
  };
  // (exclusive prefix over the slot, total) of a per-thread count
  auto slot_scan = [&](int c, int& total) -> int {
    int inc = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(kFull, inc, o); if (lane >= o) inc += u; }
    int base = 0;
    total = __shfl_sync(kFull, inc, 31);
    if (G > 1) {
      xsel ^= 1;
      if (lane == 31) xch[slot][xsel][s.warp] = inc;
      s.sync();
      total = 0;
#pragma unroll
      for (int g = 0; g < G; ++g) { const int v = xch[slot][xsel][g]; if (g < s.warp) base += v; total += v; }
    }
    return base + inc - c;
  };

  // Work items are SEGMENTS of backtests (A.seg decisions): a slot that has run a segment saves the backtest's state and
  // appends it to the ready queue again, then takes the backtest at the head of the queue.  Whole backtests as items leave
  // the last round of slots partly empty (4096 backtests on 1184 slots: 3.46 rounds cost 4); with segments in FIFO order
  // every backtest advances at the same rate and the slots run dry together.  The queue is a ring of B ids (a backtest is
  // in it at most once): push = reserve a position with an atomic increment of the tail, then publish the id; pop = advance
  // the head by compare-and-swap while head < tail, then take the id of the reserved position (it is published a moment
  // after the reservation at the latest).  An idle slot only READS head and tail until there is work.
  // done_counter counts the backtests that have left this kernel for good (finished, or suspended for the full-width kernel).
  // returns 1: a backtest is loaded; 0: none ready right now (the others are running); -1: every backtest has left the kernel
  auto fetch = [&]() -> int {
    int got = -1;
    if (tid == 0) {
      volatile int* head = A.queue_ctr;
      volatile int* tail = A.queue_ctr + 1;
#pragma unroll 1
      for (int tries = 0; tries < 16; ++tries) {
        if (*(volatile int*)A.done_counter >= A.B) { got = -2; break; }
        const int hd = *head;
        if (hd >= *tail) break;                        // empty
        if (atomicCAS(A.queue_ctr, hd, hd + 1) == hd) {
          int* cell = A.ready_ring + (unsigned)hd % (unsigned)A.B;
          int v;
          unsigned spins = 0;
          while ((v = atomicExch(cell, -1)) < 0) {     // reserved by its pusher an instant ago
            if (++spins > (1u << 24)) { printf("kmpc active-set kernel: ready queue cell never published\n"); __trap(); }   // fail loudly, never hang
          }
          got = v;
          break;
        }
      }
    }
    got = slot_bcast0(got);
    if (got == -2) return -1;
    if (got < 0) return 0;
    b = got;
    __threadfence();                                 // the state below was written by the backtest's previous slot
    const double* S = A.state + (size_t)b * A.state_ld;
#pragma unroll
    for (int q = 0; q < NQ; ++q) { const int i = tid + NT * q; if (i < N) wfull[i] = __ldcg(S + i); }
    t = (int)__ldcg(S + N + 14);
    if (tid == 0) {
      double bk[14];
#pragma unroll
      for (int i = 0; i < 14; ++i) bk[i] = __ldcg(S + N + i);
      book_load(books[slot], bk);
    }
    rb = (size_t)(A.realized_index ? A.realized_index[b] : b) * A.realized_stride;
    lam_b = A.lam ? A.lam[b] : A.lam0; tau_b = A.tau ? A.tau[b] : A.tau0;
    seg_left = (A.seg > 0) ? A.seg : 0x7fffffff;
    slot_sync();
    return 1;
  };
  // gives the backtest up: status 1 = ready for its next segment, 2 = for the full-width kernel (both with the state saved),
  // 3 = finished
  auto release = [&](int status) {
    slot_sync();
    if (status != 3) {
      double* S = A.state + (size_t)b * A.state_ld;
#pragma unroll
      for (int q = 0; q < NQ; ++q) { const int i = tid + NT * q; if (i < N) S[i] = wfull[i]; }
      if (tid == 0) book_save(books[slot], t, S + N);
    }
    __threadfence();
    slot_sync();
    if (tid == 0) {
      if (status == 1) {
        const int pos = atomicAdd(A.queue_ctr + 1, 1);
        atomicExch(A.ready_ring + (unsigned)pos % (unsigned)A.B, b);
      } else {
        A.bt_status[b] = status;
        atomicAdd(A.done_counter, 1);
      }
    }
  };
  // my thread's problem data for the current S, then the starting point
  auto start_solve = [&]() -> int {
    slot_sync();
    s.valid = tid < count;
    a = s.valid ? sid[tid] : 0;
    wc = s.valid ? wfull[a] : 0.0;
    const float y_next = (s.valid && t + 1 < A.rows) ? rnext[a] : 0.0f;
    e_next = s.load_returns(ysrc, (size_t)N, y_next, a);                                  // mpc.py:55
    return s.begin(wc, count, lam_b, tau_b, false, opt, false);
  };
  // adds the assets flagged in `add` (bit q of my thread) to S.  must = true: all of them, false if S would exceed the slot;
  // must = false (candidates): as many as fit — an asset left out is still subject to the optimality check afterwards
  auto grow = [&](unsigned add, bool must) -> bool {
    if (G == 1) {
      int base = count;
#pragma unroll
      for (int q = 0; q < NQ; ++q) base += __popc(__ballot_sync(kFull, (add >> q) & 1u));
      if (must && base > 32) return false;
      base = count;
#pragma unroll 1
      for (int q = 0; q < NQ; ++q) {
        const unsigned bal = __ballot_sync(kFull, (add >> q) & 1u);
        const int pos = base + __popc(bal & lt_mask);
        if (((add >> q) & 1u) && pos < 32) { sid[pos] = lane + 32 * q; member |= 1u << q; }
        base += __popc(bal);
      }
      count = base < 32 ? base : 32;
      return true;
    }
    int total;
    int pos = count + slot_scan(__popc(add & ((1u << NQ) - 1u)), total);
    if (must && count + total > NT) return false;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      if ((add >> q) & 1u) {
        if (pos < NT) { sid[pos] = tid + NT * q; member |= 1u << q; }
        ++pos;
      }
    }
    count = (count + total < NT) ? count + total : NT;
    return true;
  };

  int have = fetch();
  bool active = have == 1, finished = have < 0;    // active: a backtest is loaded; finished: nothing left for this kernel
  int need_start = 1;                              // 1: a new decision (choose S), 2: the same decision on a grown S, 0: iterating
  unsigned idle_trips = 0;
  unsigned pending = 0, cand = 0;                  // assets that join S at the next start / candidates (bit q of my thread)
  int st = -1;
  __syncthreads();
#pragma unroll 1
  for (unsigned trip = 0;; ++trip) {
    if (uni(!active && !finished)) {               // idle: every ready backtest was taken a moment ago; look again
      have = fetch();
      active = have == 1; finished = have < 0;
      need_start = 1; st = -1;
      if (!active && !finished) {
        __nanosleep(500);                            // nothing to do until a neighbour releases a backtest
        if (++idle_trips > (1u << 28)) {             // minutes of idling: trap instead of hanging the GPU
          if (tid == 0) printf("kmpc active-set kernel: slot idle for 2^28 trips with unfinished backtests\n");
          __trap();
        }
      }
      if (active) idle_trips = 0;
    }
    if (uni(active)) {
#pragma unroll 1
      for (;;) {
        if (uni(need_start != 0)) {
          bool bad = false;                          // a forecast outside the range the full solver accepts (screening of begin())
          if (uni(need_start == 1)) {
            // ---- a new decision: choose the active set ------------------------------------------------------------------
            yrow = A.yhat + (size_t)(A.yhat_index ? A.yhat_index[b] : b) * A.yhat_stride + (size_t)t * H * N;
            {
              const float* rrow = A.realized + rb + (size_t)(t + 1) * N;
              const bool market = (t + 1 < A.rows);
              if (staged) {
                __syncwarp();
#pragma unroll 4
                for (int i = lane; i < (H + 1) * N; i += 32)
                  ystage[i] = (i < H * N) ? yrow[i] : (market ? rrow[i - H * N] : 0.0f);
                __syncwarp();
              } else {
                ysrc = yrow; rnext = rrow;
                slot_sync();                           // the weights the previous decision left in wfull (written per asset)
              }
            }
#pragma unroll 4
            for (int i = tid; i < H * N; i += NT)
              if (!(fabsf(ysrc[i]) < 80.0f)) bad = true;               // exp() may leave the positive normal floats: looked at below
            pending = 0;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              const int i = tid + NT * q;
              if (i < N) { if (wfull[i] > kHeldThr) pending |= 1u << q; else wfull[i] = 0.0; }
            }
            cand = 0;
            const int n_stage = (opt.active_set == 2) ? 0 : H;    // 2 (test hook): held assets only, the repair path does the rest
#pragma unroll 1
            for (int k = 0; k < n_stage; ++k) {
              float v[NQ];
#pragma unroll
              for (int q = 0; q < NQ; ++q) {
                const int i = tid + NT * q;
                v[q] = (i < N) ? ysrc[k * N + i] : -CUDART_INF_F;
              }
#pragma unroll 1
              for (int rep = 0; rep < 2; ++rep) {                       // the two best forecasts of the stage
                float m = v[0];
#pragma unroll
                for (int q = 1; q < NQ; ++q) m = fmaxf(m, v[q]);
                const float wm = slot_maxf(m);
                // (G > 1: a tie between two warps makes both of their assets candidates, which is as good)
                const unsigned bal = __ballot_sync(kFull, m == wm && m > -CUDART_INF_F);
                if (bal && lane == __ffs(bal) - 1) {
                  bool done = false;
#pragma unroll
                  for (int q = 0; q < NQ; ++q) if (!done && v[q] == wm) { cand |= 1u << q; v[q] = -CUDART_INF_F; done = true; }
                }
              }
            }
            member = 0; count = 0; extra_it = 0;
          }
          const bool fits = grow(pending, true);     // the held assets (or the assets that failed the check) must all fit
          if (uni(fits && need_start == 1)) grow(cand & ~member, false);
          if (uni(!fits)) {                          // more active assets than threads: the full-width kernel takes over
            release(2);
            need_start = 1; st = -1;
            have = fetch();
            active = have == 1; finished = have < 0;
            if (!active) break;
            continue;
          }
          st = start_solve();
          if (uni(slot_any(bad))) {
            // a suspicious forecast somewhere in the universe: would the full solver's screening refuse the decision?
            bool refuse = false;
#pragma unroll 1
            for (int i = tid; i < H * N; i += NT) {
              const float r = __double2float_rn(exp((double)ysrc[i]));
              if (!(isfinite(r) && r > 0.0f)) refuse = true;
            }
            if (uni(slot_any(refuse))) {             // hold the weights, as the full solver does
#pragma unroll
              for (int k = 0; k < H; ++k) s.w[k] = wc;
              s.it_ = 0;
              st = ST_NONFINITE;
            }
          }
          need_start = 0;
        } else if (uni(st == ST_RESTART)) {                            // second attempt: same returns, same weights
          s.sync();
          st = s.begin(wc, count, lam_b, tau_b, false, opt, true);
        }
        if (uni(st == -1)) st = s.check(opt);
        if (uni(st == ST_RESTART)) continue;
        if (uni(st < 0)) break;                                        // take a Newton step
        if (uni(st <= ST_INACCURATE)) {
          // ---- converged on S: optimality conditions of the excluded assets ---------------------------------------------
          double nuk[H], irho[H], ck[H], thr[H];
#pragma unroll
          for (int k = 0; k < H; ++k) {
            nuk[k] = s.U(Ipm::U_NU, k); irho[k] = s.U(Ipm::U_IRHO, k);
            ck[k] = (s.hu() ? s.lam : 0.0) + (s.hc() ? s.U(Ipm::U_ZC, k) : 0.0);
            // g_k = nu_k - R_ik / rho_k > 0  <=>  yhat_ik < log(nu_k rho_k); log x >= 1 - 1/x spares the logarithm (nu rho is
            // within a few per cent of 1), 1e-6 covers the float32 rounding of R
            thr[k] = 1.0 - rcp_fast(nuk[k]) * irho[k] - 1e-6;
          }
          unsigned viol = 0;
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const int i = tid + NT * q;
            if (i < N && !((member >> q) & 1u)) {
              float yv[H];
              bool safe = true;
#pragma unroll
              for (int k = 0; k < H; ++k) { yv[k] = ysrc[k * N + i]; safe = safe && ((double)yv[k] < thr[k]); }
              if (!safe) {
                double yk = 0.0;
                bool out = false;
#pragma unroll 1
                for (int k = H - 1; k >= 0; --k) {
                  const double Rk = (double)__double2float_rn(exp((double)yv[k]));
                  const double g = nuk[k] - Rk * irho[k];
                  yk = fmax(-ck[k], yk - g);
                  if (!(yk <= ck[k] + kVerifyTol)) out = true;
                }
                if (out) viol |= 1u << q;
              }
            }
          }
#ifdef KMPC_AS_DEBUG
          { const int anyv = __any_sync(kFull, viol != 0);
            if (lane == 0 && blockIdx.x == 0) printf("AS slot %d b %d t %d st %d it %d viol %d\n", slot, b, t, st, s.it_, anyv); }
#endif
          if (uni(slot_any(viol != 0))) {                               // somebody wants in: solve again on the larger set
            extra_it += s.it_;
            pending = viol;
            need_start = 2; st = -1;
            continue;
          }
        }
        // ---- the decision is made: portfolio step (backtest.py:175-217), as in backtest_lane_kernel -------------------
        const bool market = (t + 1 < A.rows);
        double turnover, port_ret;
        wc = portfolio_step(s, opt, wc, e_next, market, turnover, port_ret);
        if (s.valid) wfull[a] = wc;
        t += A.rebalance_freq;
        const bool last = (t >= A.n_steps);
        if (tid == 0) {
          book_decision(books[slot], A, b, st, s.it_ + extra_it, turnover, port_ret, market, last);
        }
        need_start = 1; st = -1;
        if (uni(last)) {
          slot_sync();
          if (A.final_weights) {
#pragma unroll
            for (int q = 0; q < NQ; ++q) { const int i = tid + NT * q; if (i < N) A.final_weights[(size_t)b * N + i] = wfull[i]; }
          }
        }
        if (uni(last || --seg_left == 0)) {          // the backtest is finished, or its segment is: take the next ready one
          release(last ? 3 : 1);
          have = fetch();
          active = have == 1; finished = have < 0;
          if (!active) break;
        }
      }
    }
#ifdef KMPC_AS_DEBUG
    if (trip > 30000u) {
      if (lane == 0) printf("A kernel stuck: block %d slot %d active %d b %d t %d st %d need_start %d it %d count %d\n",
                            (int)blockIdx.x, slot, (int)active, b, t, st, (int)need_start, s.it_, count);
      break;
    }
#endif
    if ((trip % KMPC_ACTIVE_SYNC_EVERY) == 0) {
      if (__syncthreads_and(finished)) break;
    }
    const bool act_u = uni(active);
    bool ok = false;
    if (act_u) {
      s.factor_a();
      ok = s.factor_b();
    }
    if (ok) {
#pragma unroll 1
      for (int phase = 0; phase < 2; ++phase) s.newton_phase(phase, opt);
    }
  }
}

#ifndef KMPC_LANE_SLOT_WARPS
#define KMPC_LANE_SLOT_WARPS 8      // warps per block = slots per block x G
#endif
template <int G> struct LaneSlots { static constexpr int P = (KMPC_LANE_SLOT_WARPS / G) < 1 ? 1 : (KMPC_LANE_SLOT_WARPS / G); };

template <typename K>
static int lane_blocks_per_sm(K kernel, int threads, size_t smem) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem);
  return nb < 1 ? 1 : nb;
}

// which kernel(s) to launch: 1 = FIX only, 0 = generic only, 2 = both, gated on the device flag
static int lane_fix_plan(const double* lam, const double* tau, double lam0, double tau0, int allow_short, double dual_init,
                         int n, int* flag, cudaStream_t st) {
  if (allow_short || !(dual_init > 0.0)) return 0;
  if (!lam && !tau) return (lam0 > 0.0 && tau0 > 0.0) ? 1 : 0;
  lane_flags_kernel<<<1, 256, 0, st>>>(lam, tau, lam0, tau0, n, flag);
  return 2;
}

template <int H, int G>
static int launch_mpc_lane(const MpcSolveArgs& A, int sm_count, cudaStream_t st) {
  const size_t smem = (size_t)LaneIpm<H, G, (G > 4 || H > 5), false>::SMEM_DOUBLES * sizeof(double);
  static PerDeviceInt t0, t1;          // blocks per SM (and the shared-memory attribute) of the two instantiations
  const int bps0 = t0.get([&] { return lane_blocks_per_sm(mpc_solve_lane_kernel<H, G, false>, 32 * G, smem); });
  const int bps1 = t1.get([&] { return lane_blocks_per_sm(mpc_solve_lane_kernel<H, G, true>, 32 * G, smem); });
  const int plan = lane_fix_plan(A.lam, A.tau, A.lam0, A.tau0, A.allow_short, A.opt.dual_init, A.P, A.fix_flag, st);
  auto nblocks = [&](int bps) { int b = A.P < sm_count * bps ? A.P : sm_count * bps; return b < 1 ? 1 : b; };
  if (plan != 0) mpc_solve_lane_kernel<H, G, true><<<nblocks(bps1), 32 * G, smem, st>>>(A, plan == 2 ? 1 : -1);
  if (plan != 1) mpc_solve_lane_kernel<H, G, false><<<nblocks(bps0), 32 * G, smem, st>>>(A, plan == 2 ? 0 : -1);
  return (int)cudaGetLastError();
}
template <int H, int G>
static int launch_bt_lane(const BacktestArgs& A, int sm_count, cudaStream_t st) {
  constexpr int P = LaneSlots<G>::P;
  const size_t smem = (size_t)P * LaneIpm<H, G, (G > 4 || H > 5), false>::SMEM_DOUBLES * sizeof(double);
  static PerDeviceInt t0, t1;
  const int bps0 = t0.get([&] { return lane_blocks_per_sm(backtest_lane_kernel<H, G, P, false>, 32 * G * P, smem); });
  const int bps1 = t1.get([&] { return lane_blocks_per_sm(backtest_lane_kernel<H, G, P, true>, 32 * G * P, smem); });
  const int plan = lane_fix_plan(A.lam, A.tau, A.lam0, A.tau0, A.allow_short, A.opt.dual_init, A.B, A.fix_flag, st);
  const int want = (A.B + P - 1) / P;
  auto nblocks = [&](int bps) { int b = want < sm_count * bps ? want : sm_count * bps; return b < 1 ? 1 : b; };
  if (plan != 0) backtest_lane_kernel<H, G, P, true><<<nblocks(bps1), 32 * G * P, smem, st>>>(A, plan == 2 ? 1 : -1);
  if (plan != 1) backtest_lane_kernel<H, G, P, false><<<nblocks(bps0), 32 * G * P, smem, st>>>(A, plan == 2 ? 0 : -1);
  return (int)cudaGetLastError();
}

// Active-set pipeline for problems of G > 1 warps (32 < N <= 512), three launches on one stream (no host synchronisation):
//   phase 1  backtest_lane_kernel<H, G>: every backtest from the equal-weight start until few assets are held;
//   active   backtest_active_kernel<H>: reduced solves, one warp per problem;
//   phase 3  backtest_lane_kernel<H, G>: backtests the active-set kernel suspended (active set beyond 32 assets) run
//            full-width until few assets are held again (at least one decision), then return to the ready queue;
//   active   the reduced-solve kernel once more for those;
//   phase 2  backtest_lane_kernel<H, G>: whatever was suspended a second time, to its end.
// With nothing suspended the last three launches find no work (a few microseconds each).
template <int H>
static int launch_bt_active(const BacktestArgs& A, int sm_count, cudaStream_t st) {
  constexpr int P = ActiveSlots<H>::P;
  using Ipm = LaneIpm<H, 1, KMPC_ACTIVE_LOC(H), false>;
  const int plan = lane_fix_plan(A.lam, A.tau, A.lam0, A.tau0, A.allow_short, A.opt.dual_init, A.B, A.fix_flag, st);
  const int want = (A.B + P - 1) / P;
  auto go = [&](auto nq) {
    constexpr int NQ = decltype(nq)::value;
    // per slot: the solver's slice, the weights of the whole universe, the lane -> asset table and the staged forecasts
    const size_t fixed = (size_t)P * (Ipm::SMEM_DOUBLES + 32 * NQ) * sizeof(double) + (size_t)P * 32 * sizeof(int);
    const size_t smem_max = fixed + (size_t)P * (H + 1) * 32 * NQ * sizeof(float);
    const size_t smem = fixed + (size_t)P * (H + 1) * A.N * sizeof(float);
    static PerDeviceInt t0, t1;
    const int bps0 = t0.get([&] { return lane_blocks_per_sm(backtest_active_kernel<H, P, false, NQ>, 32 * P, smem_max); });
    const int bps1 = t1.get([&] { return lane_blocks_per_sm(backtest_active_kernel<H, P, true, NQ>, 32 * P, smem_max); });
    auto nblocks = [&](int bps) { int b = want < sm_count * bps ? want : sm_count * bps; return b < 1 ? 1 : b; };
    if (plan != 0) backtest_active_kernel<H, P, true, NQ><<<nblocks(bps1), 32 * P, smem, st>>>(A, plan == 2 ? 1 : -1);
    if (plan != 1) backtest_active_kernel<H, P, false, NQ><<<nblocks(bps0), 32 * P, smem, st>>>(A, plan == 2 ? 0 : -1);
  };
  if (A.N <= 64) go(std::integral_constant<int, 2>{});
  else if (A.N <= 128) go(std::integral_constant<int, 4>{});
  else return -2;                                   // larger universes take the wide kernel (launch_bt_active_wide)
  return (int)cudaGetLastError();
}

// The reduced-solve kernel for universes of 129..512 assets: KMPC_WIDE_G warps per problem (up to 32 G active assets), forecasts
// read from global memory / L2.  Compiled with the (H, 4) variants of the horizons that have a 16-warp full-width kernel.
// Measured per Newton iteration and problem at H = 10, one problem per SM (scripts/lane_latency.py): 57-59 us for 32, 64 and
// 128 assets on 1, 2, 4 warps; 197 / 215 us for 256 / 500 assets on the 16 warps of the full-width kernel.
#ifndef KMPC_WIDE_G
#define KMPC_WIDE_G 8     // config 3 (148 backtests x 500 assets, H = 10): 4 warps 516 ms, 8 warps 371 ms per pass (one-warp route: 1010 ms)
#endif
constexpr int kWideG = KMPC_WIDE_G, kWideNQ = 16 / KMPC_WIDE_G;
template <int H>
static int launch_bt_active_wide(const BacktestArgs& A, int sm_count, cudaStream_t st) {
  constexpr int G = kWideG, NQ = kWideNQ, P = ActiveSlots<H, G>::P;
  using Ipm = LaneIpm<H, G, ActiveSlots<H, G>::LOC, false>;
  if (A.N > 32 * G * NQ) return -2;
  const int plan = lane_fix_plan(A.lam, A.tau, A.lam0, A.tau0, A.allow_short, A.opt.dual_init, A.B, A.fix_flag, st);
  const int want = (A.B + P - 1) / P;
  const size_t smem = (size_t)P * (Ipm::SMEM_DOUBLES + 32 * G * NQ) * sizeof(double) + (size_t)P * 32 * G * sizeof(int);
  static PerDeviceInt t0, t1;
  const int bps0 = t0.get([&] { return lane_blocks_per_sm(backtest_active_kernel<H, P, false, NQ, G>, 32 * G * P, smem); });
  const int bps1 = t1.get([&] { return lane_blocks_per_sm(backtest_active_kernel<H, P, true, NQ, G>, 32 * G * P, smem); });
  auto nblocks = [&](int bps) { int b = want < sm_count * bps ? want : sm_count * bps; return b < 1 ? 1 : b; };
  if (plan != 0) backtest_active_kernel<H, P, true, NQ, G><<<nblocks(bps1), 32 * G * P, smem, st>>>(A, plan == 2 ? 1 : -1);
  if (plan != 1) backtest_active_kernel<H, P, false, NQ, G><<<nblocks(bps0), 32 * G * P, smem, st>>>(A, plan == 2 ? 0 : -1);
  return (int)cudaGetLastError();
}

}  // namespace kmpc
