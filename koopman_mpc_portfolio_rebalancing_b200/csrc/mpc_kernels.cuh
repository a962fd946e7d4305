// Kernels around the warp IPM solver:
//   mpc_solve_kernel   P independent MPC problems (drop-in for mpc.solve_mpc_log_utility, mpc.py:27-117)
//   backtest_kernel    the whole rebalancing loop of run_backtest (backtest.py:173-217) for many independent
//                      backtests, batch-resident: per step  MPC solve -> turnover/cost/value -> realised
//                      return -> drift, with the metrics of calculate_metrics (backtest.py:221-249)
//                      accumulated on the fly.
// One warp per problem / per backtest; lanes own assets lane, lane+32, ...
#pragma once
#include "kmpc_internal.cuh"
#include "mpc_ipm.cuh"

namespace kmpc {

// round_f32(exp_f64(y)): the platform-independent stand-in for numpy's fp32 exp (mpc.py:55, backtest.py:193)
__device__ __forceinline__ float exp_cr32(float y) { return __double2float_rn(exp((double)y)); }

constexpr int kWarpsPerBlock = 1;   // one problem per warp; ~50 KB of shared memory per warp at H=5, N<=64

template <int H, int APT, int NS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 1)
mpc_solve_kernel(MpcSolveArgs A) {
  using Ipm = WarpIpm<H, APT, NS>;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int wid = blockIdx.x * kWarpsPerBlock + wib, nwarps = gridDim.x * kWarpsPerBlock;
  Ipm s;
  s.bind(smem + (size_t)wib * Ipm::SMEM_DOUBLES, lane, A.N);
  const int N = A.N;
  const IpmOptions opt = A.opt;
  for (int p = wid; p < A.P; p += nwarps) {
    double w0[APT];
#pragma unroll
    for (int a = 0; a < APT; ++a) {
      const int i = lane + 32 * a;
            w0[a] = s.ok(a) ? A.w_cur[(size_t)p * N + i] : 0.0;
#pragma unroll
      for (int k = 0; k < H; ++k) {
        double r = 1.0;
        if (s.ok(a)) {
          if (A.yhat) r = (double)exp_cr32(A.yhat[((size_t)p * H + k) * N + i]);
          else r = exp(A.yhat64[((size_t)p * H + k) * N + i]);
        }
        if (s.ok(a)) s.F(Ipm::RR, k, a) = r;
      }
    }
    const double lam = A.lam ? A.lam[p] : A.lam0;
    const double tau = A.tau ? A.tau[p] : A.tau0;
    int iters; double kkt[3];
    const int st = s.solve(w0, N, lam, tau, A.allow_short != 0, opt, iters, kkt);
    const double val = (st <= ST_INACCURATE) ? s.objective(w0) : CUDART_NAN;
#pragma unroll
    for (int a = 0; a < APT; ++a)
      if (s.ok(a)) {
#pragma unroll
        for (int k = 0; k < H; ++k) A.w_out[((size_t)p * H + k) * N + lane + 32 * a] = s.F(Ipm::WW, k, a);
      }
    if (lane == 0) {
      if (A.obj) A.obj[p] = val;
      if (A.kkt) { A.kkt[3 * p] = kkt[0]; A.kkt[3 * p + 1] = kkt[1]; A.kkt[3 * p + 2] = kkt[2]; }
      if (A.status) A.status[p] = st;
      if (A.iters) A.iters[p] = iters;
    }
    __syncwarp();
  }
}

template <int H, int APT, int NS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 1)
backtest_kernel(BacktestArgs A) {
  using Ipm = WarpIpm<H, APT, NS>;
  extern __shared__ double smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  Ipm s;
  s.bind(smem + (size_t)wib * Ipm::SMEM_DOUBLES, lane, A.N);
  const int N = A.N;
  const IpmOptions opt = A.opt;
  __shared__ int next_b[kWarpsPerBlock];
  for (;;) {
    // dynamic work distribution: backtests differ in iteration counts
    if (lane == 0) next_b[wib] = atomicAdd(A.work_counter, 1);
    __syncwarp();
    const int b = next_b[wib];
    __syncwarp();
    if (b >= A.B) break;
    const size_t yb = (size_t)(A.yhat_index ? A.yhat_index[b] : b) * A.yhat_stride;
    const size_t rb = (size_t)(A.realized_index ? A.realized_index[b] : b) * A.realized_stride;
    const double lam = A.lam ? A.lam[b] : A.lam0;
    const double tau = A.tau ? A.tau[b] : A.tau0;
    const double ccoef = A.cost_coeff ? A.cost_coeff[b] : A.cost_coeff0;
    double V = A.capital ? A.capital[b] : A.capital0;
    double wc[APT];
#pragma unroll
    for (int a = 0; a < APT; ++a) {
            wc[a] = s.ok(a) ? 1.0 / (double)N : 0.0;          // backtest.py:161
    }
    // running metric state (backtest.py:221-249)
    double mean = 0.0, m2 = 0.0, cum = 1.0, peak = -CUDART_INF, maxdd = CUDART_INF, sum_turn = 0.0, v_first = 0.0;
    int n = 0, n_opt = 0, n_inacc = 0, n_fail = 0;
    long long it_total = 0;
    for (int t = 0; t < A.n_steps; t += A.rebalance_freq) {
      // ---- forecast of this step -> gross returns (mpc.py:55) ----
      const float* yh = A.yhat + yb + (size_t)t * H * N;
#pragma unroll
      for (int a = 0; a < APT; ++a)
#pragma unroll
        for (int k = 0; k < H; ++k)
          if (s.ok(a)) s.F(Ipm::RR, k, a) = (double)exp_cr32(yh[k * N + lane + 32 * a]);
      int iters; double kkt[3];
      const int st = s.solve(wc, N, lam, tau, A.allow_short != 0, opt, iters, kkt);
      it_total += iters;
      n_opt += (st == ST_OPTIMAL); n_inacc += (st == ST_INACCURATE); n_fail += (st >= ST_FAILED);
      // ---- apply first-stage weights (backtest.py:131), costs (179-184) ----
      double tn = 0.0;
#pragma unroll
      for (int a = 0; a < APT; ++a)
        if (s.ok(a)) tn += fabs(s.F(Ipm::WW, 0, a) - wc[a]);
      const double turnover = warp_sum(tn);
      const double cost = ccoef * turnover * V;
      V -= cost;
      // ---- market step (backtest.py:187-208) ----
      double port_ret = 0.0;
      if (t + 1 < A.rows) {
        const float* rr = A.realized + rb + (size_t)(t + 1) * N;
        float r32[APT];
        double pr = 0.0;
#pragma unroll
        for (int a = 0; a < APT; ++a) {
          r32[a] = s.ok(a) ? __fsub_rn(exp_cr32(rr[lane + 32 * a]), 1.0f) : 0.0f;   // f32, backtest.py:193
          if (s.ok(a)) pr += s.F(Ipm::WW, 0, a) * (double)r32[a];
        }
        port_ret = warp_sum(pr);
        V *= (1.0 + port_ret);
        double denom = 1.0 + port_ret;
        if (fabs(denom) < 1e-8) denom = 1e-8;
#pragma unroll
        for (int a = 0; a < APT; ++a) {
          wc[a] = 0.0;
          if (s.ok(a)) wc[a] = s.F(Ipm::WW, 0, a) * (double)__fadd_rn(1.0f, r32[a]) / denom;   // (1.0 + f32) stays f32
        }
      } else {
#pragma unroll
        for (int a = 0; a < APT; ++a) { wc[a] = 0.0; if (s.ok(a)) wc[a] = s.F(Ipm::WW, 0, a); }
      }
      // ---- history row + metric accumulators ----
      if (A.history && lane == 0) {
        double* hrow = A.history + ((size_t)b * A.n_hist + n) * 4;
        hrow[0] = V; hrow[1] = port_ret; hrow[2] = turnover; hrow[3] = cost;
      }
      if (n == 0) v_first = V;
      ++n;
      const double dlt = port_ret - mean;
      mean += dlt / (double)n;
      m2 += dlt * (port_ret - mean);
      cum *= (1.0 + port_ret);
      peak = fmax(peak, cum);
      maxdd = fmin(maxdd, (cum - peak) / peak);
      sum_turn += turnover;
    }
    if (lane == 0) {
      double* m = A.metrics + (size_t)b * 5;
      if (n > 0) {
        const double sd = sqrt(m2 / (double)n);
        m[0] = sqrt(252.0) * mean / (sd + 1e-8);
        m[1] = maxdd;
        m[2] = sum_turn / (double)n;
        m[3] = V;
        m[4] = V / v_first - 1.0;
      } else { m[0] = m[1] = m[2] = m[3] = m[4] = CUDART_NAN; }
      if (A.solve_stats) {
        long long* ss = A.solve_stats + (size_t)b * 4;
        ss[0] = n_opt; ss[1] = n_inacc; ss[2] = n_fail; ss[3] = it_total;
      }
    }
    if (A.final_weights) {
#pragma unroll
      for (int a = 0; a < APT; ++a)
        if (s.ok(a)) A.final_weights[(size_t)b * N + lane + 32 * a] = wc[a];
    }
  }
}

template <typename K>
static int blocks_per_sm_for(K kernel, size_t smem) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kWarpsPerBlock * 32, smem);
  return nb < 1 ? 1 : nb;
}

template <int H, int APT, int NS>
static int launch_mpc(const MpcSolveArgs& A, int sm_count, cudaStream_t st) {
  const size_t smem = (size_t)kWarpsPerBlock * WarpIpm<H, APT, NS>::SMEM_DOUBLES * sizeof(double);
  static const int bps = blocks_per_sm_for(mpc_solve_kernel<H, APT, NS>, smem);
  int blocks = (A.P + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int cap = sm_count * bps;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  mpc_solve_kernel<H, APT, NS><<<blocks, kWarpsPerBlock * 32, smem, st>>>(A);
  return (int)cudaGetLastError();
}
template <int H, int APT, int NS>
static int launch_bt(const BacktestArgs& A, int sm_count, cudaStream_t st) {
  const size_t smem = (size_t)kWarpsPerBlock * WarpIpm<H, APT, NS>::SMEM_DOUBLES * sizeof(double);
  static const int bps = blocks_per_sm_for(backtest_kernel<H, APT, NS>, smem);
  int blocks = (A.B + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int cap = sm_count * bps;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  backtest_kernel<H, APT, NS><<<blocks, kWarpsPerBlock * 32, smem, st>>>(A);
  return (int)cudaGetLastError();
}

}  // namespace kmpc
