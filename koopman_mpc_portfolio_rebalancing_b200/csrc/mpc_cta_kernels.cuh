// Kernels around the CTA-per-problem IPM solver (mpc_cta.cuh): same entry points and semantics as
// mpc_kernels.cuh (mpc_solve = mpc.py:27-117, backtest = backtest.py:173-249), one thread block per problem /
// per backtest, thread (k, i) = (stage, asset).
#pragma once
#include "kmpc_internal.cuh"
#include "mpc_cta.cuh"

#ifndef KMPC_CTA_MINB
#define KMPC_CTA_MINB 2      // resident blocks per SM the register allocation is sized for
#endif

namespace kmpc {

__device__ __forceinline__ float exp_cr32_cta(float y) { return __double2float_rn(exp((double)y)); }

template <int H, int G>
__global__ void __launch_bounds__(32 * H * G, KMPC_CTA_MINB)
mpc_solve_cta_kernel(MpcSolveArgs A) {
  using Ipm = CtaIpm<H, G>;
  extern __shared__ double smem[];
  Ipm s;
  s.bind(smem, A.N);
  const int N = A.N;
  const IpmOptions opt = A.opt;
  for (int p = blockIdx.x; p < A.P; p += gridDim.x) {
    if (s.k == 0) smem[Ipm::OFF_W0 + s.i] = s.valid ? A.w_cur[(size_t)p * N + s.i] : 0.0;
    s.R = 1.0;
    if (s.valid) {
      const size_t idx = ((size_t)p * H + s.k) * N + s.i;
      s.R = A.yhat ? (double)exp_cr32_cta(A.yhat[idx]) : exp(A.yhat64[idx]);
    }
    __syncthreads();
    const double lam = A.lam ? A.lam[p] : A.lam0;
    const double tau = A.tau ? A.tau[p] : A.tau0;
    int iters; double kkt[3];
    const int st = s.solve(N, lam, tau, A.allow_short != 0, opt, iters, kkt);
    double val = CUDART_NAN;
    if (st <= ST_INACCURATE) val = s.objective();
    if (s.valid) A.w_out[((size_t)p * H + s.k) * N + s.i] = s.w;
    if (threadIdx.x == 0) {
      if (A.obj) A.obj[p] = val;
      if (A.kkt) { A.kkt[3 * p] = kkt[0]; A.kkt[3 * p + 1] = kkt[1]; A.kkt[3 * p + 2] = kkt[2]; }
      if (A.status) A.status[p] = st;
      if (A.iters) A.iters[p] = iters;
    }
    __syncthreads();
  }
}

template <int H, int G>
__global__ void __launch_bounds__(32 * H * G, KMPC_CTA_MINB)
backtest_cta_kernel(BacktestArgs A) {
  using Ipm = CtaIpm<H, G>;
  extern __shared__ double smem[];
  __shared__ int next_b;
  Ipm s;
  s.bind(smem, A.N);
  const int N = A.N;
  const IpmOptions opt = A.opt;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) next_b = atomicAdd(A.work_counter, 1);   // dynamic: backtests differ in iteration counts
    __syncthreads();
    const int b = next_b;
    if (b >= A.B) break;
    const size_t yb = (size_t)(A.yhat_index ? A.yhat_index[b] : b) * A.yhat_stride;
    const size_t rb = (size_t)(A.realized_index ? A.realized_index[b] : b) * A.realized_stride;
    const double lam = A.lam ? A.lam[b] : A.lam0;
    const double tau = A.tau ? A.tau[b] : A.tau0;
    const double ccoef = A.cost_coeff ? A.cost_coeff[b] : A.cost_coeff0;
    double V = A.capital ? A.capital[b] : A.capital0;
    if (s.k == 0) smem[Ipm::OFF_W0 + s.i] = s.valid ? 1.0 / (double)N : 0.0;          // backtest.py:161
    double mean = 0.0, m2 = 0.0, cum = 1.0, peak = -CUDART_INF, maxdd = CUDART_INF, sum_turn = 0.0, v_first = 0.0;
    int n = 0, n_opt = 0, n_inacc = 0, n_fail = 0;
    long long it_total = 0;
    for (int t = 0; t < A.n_steps; t += A.rebalance_freq) {
      s.R = s.valid ? (double)exp_cr32_cta(A.yhat[yb + ((size_t)t * H + s.k) * N + s.i]) : 1.0;   // mpc.py:55
      __syncthreads();
      int iters; double kkt[3];
      const int st = s.solve(N, lam, tau, A.allow_short != 0, opt, iters, kkt);
      it_total += iters;
      n_opt += (st == ST_OPTIMAL); n_inacc += (st == ST_INACCURATE); n_fail += (st >= ST_FAILED);
      // first-stage weights (backtest.py:131) are held by the stage-0 threads
      const bool s0 = (s.k == 0) && s.valid;
      const double wc = s0 ? smem[Ipm::OFF_W0 + s.i] : 0.0;
      float r32 = 0.0f;
      const bool market = (t + 1 < A.rows);
      if (s0 && market) r32 = __fsub_rn(exp_cr32_cta(A.realized[rb + (size_t)(t + 1) * N + s.i]), 1.0f);   // backtest.py:193
      double v[2] = {s0 ? fabs(s.w - wc) : 0.0, s0 ? s.w * (double)r32 : 0.0}, S[2], T[2];
      s.template reduce_sum<2>(v, S, T);
      const double turnover = T[0];
      const double cost = ccoef * turnover * V;
      V -= cost;
      double port_ret = 0.0;
      double wnew = s.w;
      if (market) {
        port_ret = T[1];
        V *= (1.0 + port_ret);
        double denom = 1.0 + port_ret;
        if (fabs(denom) < 1e-8) denom = 1e-8;
        wnew = s.w * (double)__fadd_rn(1.0f, r32) / denom;                              // (1.0 + f32) stays f32
      }
      __syncthreads();
      if (s0) smem[Ipm::OFF_W0 + s.i] = wnew;
      if (A.history && threadIdx.x == 0) {
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
    __syncthreads();
    if (threadIdx.x == 0) {
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
    if (A.final_weights && s.k == 0 && s.valid) A.final_weights[(size_t)b * N + s.i] = smem[Ipm::OFF_W0 + s.i];
  }
}

template <typename K>
static int cta_blocks_per_sm(K kernel, int threads, size_t smem) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, smem);
  return nb < 1 ? 1 : nb;
}

template <int H, int G>
static int launch_mpc_cta(const MpcSolveArgs& A, int sm_count, cudaStream_t st) {
  const size_t smem = (size_t)CtaIpm<H, G>::SMEM_DOUBLES * sizeof(double);
  static const int bps = cta_blocks_per_sm(mpc_solve_cta_kernel<H, G>, 32 * H * G, smem);
  int blocks = A.P < sm_count * bps ? A.P : sm_count * bps;
  if (blocks < 1) blocks = 1;
  mpc_solve_cta_kernel<H, G><<<blocks, 32 * H * G, smem, st>>>(A);
  return (int)cudaGetLastError();
}
template <int H, int G>
static int launch_bt_cta(const BacktestArgs& A, int sm_count, cudaStream_t st) {
  const size_t smem = (size_t)CtaIpm<H, G>::SMEM_DOUBLES * sizeof(double);
  static const int bps = cta_blocks_per_sm(backtest_cta_kernel<H, G>, 32 * H * G, smem);
  int blocks = A.B < sm_count * bps ? A.B : sm_count * bps;
  if (blocks < 1) blocks = 1;
  backtest_cta_kernel<H, G><<<blocks, 32 * H * G, smem, st>>>(A);
  return (int)cudaGetLastError();
}

}  // namespace kmpc
