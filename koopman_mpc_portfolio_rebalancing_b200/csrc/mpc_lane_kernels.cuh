// Kernels around the lane-per-asset IPM solver (mpc_lane.cuh): mpc_solve = mpc.py:27-117, backtest =
// backtest.py:173-249.  mpc_solve: one block of G warps per
// problem; backtest: a persistent block hosts several backtests ("slots" of G warps each) that walk through the
// Newton iteration together.  Thread i of a slot = asset i, all stages of an asset in that thread's registers.
#pragma once
#include "kmpc_internal.cuh"
#include "mpc_lane.cuh"
#include <stdio.h>

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
  int b = 0, t = 0;
  double wc = 0.0;                                 // my asset's current weight
  float e_next = 1.0f;                             // exp(realised log-return of my asset on the day after the decision)
  auto fetch = [&]() -> bool {                     // next backtest of this slot (dynamic: iteration counts differ)
    if (s.tid == 0) next_b[slot] = atomicAdd(A.work_counter, 1);
    s.sync();
    b = __shfl_sync(kFull, next_b[slot], 0);
    if (b >= A.B) return false;
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
          if (!restart) {
            const size_t yb = (size_t)(A.yhat_index ? A.yhat_index[b] : b) * A.yhat_stride;
            const size_t rb = (size_t)(A.realized_index ? A.realized_index[b] : b) * A.realized_stride;
            const float y_next = (s.valid && t + 1 < A.rows) ? A.realized[rb + (size_t)(t + 1) * N + s.tid] : 0.0f;
            e_next = s.load_returns(A.yhat + yb + (size_t)t * H * N, (size_t)N, y_next);    // mpc.py:55
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
        double wn = s.valid ? s.w[0] : 0.0;                                                // backtest.py:131
        float r32 = 0.0f;
        if (s.valid && market) r32 = __fsub_rn(e_next, 1.0f);                              // backtest.py:193
        double v[3] = {fabs(wn - wc), wn * (double)r32, wc * (double)r32}, T[3];
        s.sync();
        s.template block_sum<3>(v, T);
        double turnover = T[0];
        double port_ret = market ? T[1] : 0.0;
        if (opt.clip_first_trade && s.tau > 0.0 && turnover > s.tau) {
          // An iterate accepted after the factorisation broke down next to the optimum (status optimal_inaccurate)
          // can sit ~1e-5 outside the turnover cap, whose slack the iteration does not re-derive from w: pull the
          // trade back onto the cap along its own direction (budget and sign constraints are kept).
          const double sc = div_fast(s.tau, turnover);
          wn = fma(sc, wn - wc, wc);
          if (market) port_ret = fma(sc, T[1] - T[2], T[2]);
          turnover = s.tau;
        }
        wc = wn;
        if (market) {
          double denom = 1.0 + port_ret;
          if (fabs(denom) < 1e-8) denom = 1e-8;
          wc = div_fast(wn * (double)__fadd_rn(1.0f, r32), denom);                         // (1.0 + f32) stays f32
        }
        t += A.rebalance_freq;
        const bool last = (t >= A.n_steps);
        if (s.tid == 0) {
          SlotBook& k = books[slot];
          k.it_total += s.it_;
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
        need_start = true; st = -1;
        if (uni(last)) {
          if (A.final_weights && s.valid) A.final_weights[(size_t)b * N + s.tid] = wc;
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

}  // namespace kmpc
