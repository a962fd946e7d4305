// Mean-variance MPC (/root/reference/mpc.py:119-184, solve_mpc_mean_variance):
//
//   max  sum_t [ w_t . mu_t - gamma w_t' Sigma w_t ] - lam sum_t ||w_t - w_{t-1}||_1          (w_0 = current weights)
//   s.t. 1'w_t = 1,  w_t >= 0 (unless allow_short);  no turnover cap (mpc.py:139-176)
//
// Same fp64 primal-dual interior point as the log-utility solver (Mehrotra predictor-corrector, split steps,
// epigraph u >= |w_t - w_{t-1}| eliminated per asset, proximal term delta, loose acceptance), but the stage cost has a
// DENSE Hessian 2 gamma Sigma, so the reduced Newton matrix  M = blockdiag(2 gamma Sigma) + T  (T = the per-asset
// tridiagonal of mpc_lane.cuh) is factorised densely: one warp per problem, M (n x n, n = H N <= 160) in the warp's
// slice of shared memory, in-place Cholesky, triangular solves for the right-hand side and the H budget columns, an
// H x H Schur complement for the budget multipliers.  Lane l owns variables l, l+32, ... (v = t N + i).
// The reference's only caller is MarkowitzStrategy (baselines.py:24-106) with H = 1, N assets: n = N.
// Oracle: oracle/mpc_oracle.py::solve_mv_dense (generic dense IPM on the explicit constraint matrix).
#include <cuda_runtime.h>
#include <math.h>
#include <math_constants.h>
#include <stdint.h>
#include "kmpc_internal.cuh"

namespace kmpc {

constexpr int MV_MAX_SLOTS = 5;            // n <= 160
constexpr int MV_MAX_H = 8;

struct MvArgs {
  const double* mu;       // [P,H,N]
  const double* sigma;    // [P,N,N] or [N,N] when sigma_stride == 0
  long long sigma_stride;
  const double* w_cur;    // [P,N]
  double gamma, lam;
  int allow_short, P, H, N;
  double* w_out;          // [P,H,N]
  double* obj;            // [P]
  double* kkt;            // [P,3]
  int* status;            // [P]
  int* iters;             // [P]
  IpmOptions opt;
};

__device__ __forceinline__ double mv_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double mv_warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// per-warp shared memory: M [n][ldm], X [(H+1)][n] (right-hand sides / solutions), vec [n] (neighbour exchange),
// S [H][H+1] (Schur system)
__device__ __forceinline__ size_t mv_smem_doubles(int n, int H) {
  const int ldm = n | 1;
  return (size_t)n * ldm + (size_t)(H + 1) * n + n + (size_t)H * (H + 1) + 8;
}

__global__ void __launch_bounds__(32, 1)
mpc_mv_kernel(MvArgs A) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x;
  const int H = A.H, N = A.N, n = H * N, ldm = n | 1;
  const int slots = (n + 31) / 32;
  double* M = smem;
  double* X = M + (size_t)n * ldm;          // X[c*n + v], c = 0: rhs, c = 1 + t: budget column of stage t
  double* vec = X + (size_t)(H + 1) * n;
  double* S = vec + n;
  const IpmOptions opt = A.opt;
  const double lam = A.lam, gamma = A.gamma;
  const bool has_u = lam > 0.0, has_w = !A.allow_short;

  for (int p = blockIdx.x; p < A.P; p += gridDim.x) {
    const double* Sig = A.sigma + (size_t)p * A.sigma_stride;
    double mu[MV_MAX_SLOTS], w0[MV_MAX_SLOTS], w[MV_MAX_SLOTS], sp[MV_MAX_SLOTS], sq[MV_MAX_SLOTS], zw[MV_MAX_SLOTS],
        zp[MV_MAX_SLOTS], zq[MV_MAX_SLOTS];
    int tt[MV_MAX_SLOTS], ii[MV_MAX_SLOTS];
    bool ok[MV_MAX_SLOTS];
    double nu[MV_MAX_H];
    bool bad = false;
#pragma unroll
    for (int s = 0; s < MV_MAX_SLOTS; ++s) {
      const int v = lane + 32 * s;
      ok[s] = (s < slots) && (v < n);
      tt[s] = ok[s] ? v / N : 0; ii[s] = ok[s] ? v - tt[s] * N : 0;
      mu[s] = ok[s] ? A.mu[(size_t)p * n + v] : 0.0;
      w0[s] = ok[s] ? A.w_cur[(size_t)p * N + ii[s]] : 0.0;
      if (ok[s] && !(isfinite(mu[s]) && isfinite(w0[s]))) bad = true;
    }
    int status = ST_FAILED, iters = 0;
    double kkt[3] = {CUDART_NAN, CUDART_NAN, CUDART_NAN};
    if (__any_sync(kFull, bad)) status = ST_NONFINITE;
    // ---- initial point (oracle _initial_point with tau = 0; dual-feasible start as in mpc_lane.cuh) -------------
    double delta = opt.delta;
    if (status != ST_NONFINITE) {
      double sb = 0.0;
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s] && tt[s] == 0) sb += A.allow_short ? w0[s] : fmax(w0[s], 0.0);
      sb = mv_warp_sum(sb);
      const double invN = 1.0 / (double)N, eps = 0.1, dlk = 0.05 * invN;
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) {
        const double base = A.allow_short ? w0[s] : fmax(w0[s], 0.0);
        const double b = (sb > 0.0) ? base / sb : invN;
        w[s] = ok[s] ? (1.0 - eps) * b + eps * invN : 1.0;
        const double d0 = (ok[s] && tt[s] == 0) ? w[s] - w0[s] : 0.0;
        const double u = fabs(d0) + dlk;
        sp[s] = has_u ? u - d0 : 1.0; sq[s] = has_u ? u + d0 : 1.0;
        zp[s] = has_u ? 0.5 * fmax(lam, opt.dual_init) : 0.0; zq[s] = zp[s];
      }
      // gradient of the smooth part at the start: -mu + 2 gamma Sigma w_t;  nu_t = max_i(-grad) + dual_init
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s]) vec[lane + 32 * s] = w[s];
      __syncwarp();
      double grad[MV_MAX_SLOTS];
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) {
        double a = 0.0;
        if (ok[s]) for (int j = 0; j < N; ++j) a = fma(Sig[(size_t)ii[s] * N + j], vec[tt[s] * N + j], a);
        grad[s] = fma(2.0 * gamma, a, -mu[s]);
      }
      for (int t = 0; t < H; ++t) {
        double mx = -CUDART_INF;
#pragma unroll
        for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s] && tt[s] == t) mx = fmax(mx, -grad[s]);
        nu[t] = mv_warp_max(mx) + (has_w ? opt.dual_init : 0.0);
      }
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) {
        double nut = 0.0;
        for (int t = 0; t < H; ++t) if (t == tt[s]) nut = nu[t];
        zw[s] = (has_w && ok[s]) ? grad[s] + nut : 0.0;
      }
      __syncwarp();
    }
    const double mcount = (has_w ? (double)n : 0.0) + (has_u ? 2.0 * n : 0.0);
    // ---- iterations ---------------------------------------------------------------------------------------------
    for (int it = 1; status == ST_FAILED && it <= opt.max_iter + 1; ++it) {
      iters = it;
      // residuals: Sigma w (dense), neighbours through vec
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s]) vec[lane + 32 * s] = w[s];
      __syncwarp();
      double grad[MV_MAX_SLOTS], y[MV_MAX_SLOTS];
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) {
        double a = 0.0;
        if (ok[s]) for (int j = 0; j < N; ++j) a = fma(Sig[(size_t)ii[s] * N + j], vec[tt[s] * N + j], a);
        grad[s] = fma(2.0 * gamma, a, -mu[s]);
        y[s] = zp[s] - zq[s];
      }
      __syncwarp();
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s]) vec[lane + 32 * s] = y[s];
      __syncwarp();
      double dres = 0.0, gap = 0.0, pres = 0.0;
      double rp[MV_MAX_H];
      for (int t = 0; t < H; ++t) {
        double sw = 0.0;
#pragma unroll
        for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s] && tt[s] == t) sw += w[s];
        rp[t] = mv_warp_sum(sw) - 1.0;
        pres = fmax(pres, fabs(rp[t]));
      }
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) {
        if (!ok[s]) continue;
        const double yn = (tt[s] + 1 < H) ? vec[lane + 32 * s + N] : 0.0;
        double nut = 0.0;
        for (int t = 0; t < H; ++t) if (t == tt[s]) nut = nu[t];
        dres = fmax(dres, fabs(grad[s] - (has_w ? zw[s] : 0.0) + y[s] - yn + nut));
        if (has_u) dres = fmax(dres, fabs(lam - zp[s] - zq[s]));
        if (has_w) gap = fma(w[s], zw[s], gap);
        if (has_u) gap = fma(sp[s], zp[s], fma(sq[s], zq[s], gap));
      }
      __syncwarp();
      dres = mv_warp_max(dres); gap = mv_warp_sum(gap);
      kkt[0] = pres; kkt[1] = dres; kkt[2] = gap;
      if (!isfinite(dres + gap)) break;
      if (pres < opt.tol && dres < opt.tol_dual && gap < opt.tol) { status = ST_OPTIMAL; break; }
      if (it == opt.max_iter + 1) break;
      const double mug = gap / fmax(mcount, 1.0);
      if (pres < opt.tol && gap < opt.tol) delta = fmax(0.3 * delta, 1e-9);
      // ---- barrier weights, M = blockdiag(2 gamma Sigma) + T ---------------------------------------------------------
      double iw[MV_MAX_SLOTS], isp[MV_MAX_SLOTS], isq[MV_MAX_SLOTS], ie[MV_MAX_SLOTS], ph[MV_MAX_SLOTS], ee[MV_MAX_SLOTS], ad[MV_MAX_SLOTS];
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) {
        iw[s] = 1.0 / w[s];
        ad[s] = (has_w ? zw[s] * iw[s] : 0.0) + delta;
        if (has_u) {
          isp[s] = 1.0 / sp[s]; isq[s] = 1.0 / sq[s];
          const double dp = zp[s] * isp[s], dq = zq[s] * isq[s];
          ie[s] = 1.0 / (dp + dq + delta);
          ph[s] = (dq - dp) * ie[s];
          ee[s] = (4.0 * dp * dq + 2.0 * delta * (dp + dq) + delta * delta) * ie[s];
        } else { isp[s] = isq[s] = ie[s] = 1.0; ph[s] = 0.0; ee[s] = 0.0; }
        if (ok[s]) vec[lane + 32 * s] = ee[s];
      }
      __syncwarp();
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) {
        if (!ok[s]) continue;
        const int v = lane + 32 * s;
        double* row = M + (size_t)v * ldm;
        for (int c = 0; c <= v; ++c) row[c] = 0.0;
        const int base = tt[s] * N;
        for (int j = 0; j <= ii[s]; ++j) row[base + j] = 2.0 * gamma * Sig[(size_t)ii[s] * N + j];
        const double en = (tt[s] + 1 < H) ? vec[v + N] : 0.0;            // edge to the next stage
        row[v] += ad[s] + ee[s] + en;
        if (tt[s] > 0) row[v - N] = -ee[s];                                // edge to the previous stage
      }
      __syncwarp();
      // ---- in-place Cholesky (lower) ---------------------------------------------------------------------------------
      bool pd = true;
      for (int j = 0; j < n; ++j) {
        const double d = M[(size_t)j * ldm + j];
        if (!(d > 0.0)) { pd = false; break; }
        const double inv = rsqrt(d);
        for (int i = j + lane; i < n; i += 32) M[(size_t)i * ldm + j] = (i == j) ? d * inv : M[(size_t)i * ldm + j] * inv;
        __syncwarp();
        for (int i = j + 1 + lane; i < n; i += 32) {
          const double lij = M[(size_t)i * ldm + j];
          double* ri = M + (size_t)i * ldm;
          for (int k = j + 1; k <= i; ++k) ri[k] = fma(-lij, M[(size_t)k * ldm + j], ri[k]);
        }
        __syncwarp();
      }
      if (!pd) break;
      // budget columns X[1+t] = M^{-1} A_t' (once per factorisation)
      auto solve_cols = [&](int c0, int c1) {          // in-place L L' solves of columns c0..c1-1 of X
        for (int j = 0; j < n; ++j) {
          const double invd = 1.0 / M[(size_t)j * ldm + j];
          for (int c = c0; c < c1; ++c) {
            double* x = X + (size_t)c * n;
            const double yj = x[j] * invd;
            __syncwarp();
            if (lane == 0) x[j] = yj;
            for (int i = j + 1 + lane; i < n; i += 32) x[i] = fma(-M[(size_t)i * ldm + j], yj, x[i]);
          }
          __syncwarp();
        }
        for (int j = n - 1; j >= 0; --j) {
          const double invd = 1.0 / M[(size_t)j * ldm + j];
          for (int c = c0; c < c1; ++c) {
            double* x = X + (size_t)c * n;
            const double xj = x[j] * invd;
            __syncwarp();
            if (lane == 0) x[j] = xj;
            for (int i = lane; i < j; i += 32) x[i] = fma(-M[(size_t)j * ldm + i], xj, x[i]);
          }
          __syncwarp();
        }
      };
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s)
        if (ok[s]) for (int t = 0; t < H; ++t) X[(size_t)(1 + t) * n + lane + 32 * s] = (t == tt[s]) ? 1.0 : 0.0;
      __syncwarp();
      solve_cols(1, 1 + H);
      // Schur matrix  S[t][t2] = A_t X2[:, t2] = sum over stage-t variables of column t2
      for (int t = 0; t < H; ++t)
        for (int t2 = 0; t2 < H; ++t2) {
          double a = 0.0;
#pragma unroll
          for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s] && tt[s] == t) a += X[(size_t)(1 + t2) * n + lane + 32 * s];
          a = mv_warp_sum(a);
          if (lane == 0) S[t * (H + 1) + t2] = a;
        }
      __syncwarp();
      // ---- predictor / corrector ------------------------------------------------------------------------------------
      double cw[MV_MAX_SLOTS], cp[MV_MAX_SLOTS], cq[MV_MAX_SLOTS];
      double dw[MV_MAX_SLOTS], dsp[MV_MAX_SLOTS], dsq[MV_MAX_SLOTS], dzw[MV_MAX_SLOTS], dzp[MV_MAX_SLOTS], dzq[MV_MAX_SLOTS];
      double dnu[MV_MAX_H];
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) { cw[s] = cp[s] = cq[s] = 0.0; }
      double aa = 1.0, ab = 1.0;
      for (int phase = (mcount > 0.0 ? 0 : 1); phase < 2; ++phase) {
        // right-hand side r = g_w - Delta'(ph g_u)
        double gu[MV_MAX_SLOTS], tq[MV_MAX_SLOTS], pg[MV_MAX_SLOTS];
#pragma unroll
        for (int s = 0; s < MV_MAX_SLOTS; ++s) {
          const double a1 = has_u ? cp[s] * isp[s] : 0.0, a2 = has_u ? cq[s] * isq[s] : 0.0;
          tq[s] = a1 - a2;
          gu[s] = has_u ? -lam + a1 + a2 : 0.0;
          pg[s] = ph[s] * gu[s];
          if (ok[s]) vec[lane + 32 * s] = tq[s] + pg[s];          // both enter as  - x_t + x_{t+1}
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < MV_MAX_SLOTS; ++s) {
          if (!ok[s]) continue;
          double nut = 0.0;
          for (int t = 0; t < H; ++t) if (t == tt[s]) nut = nu[t];
          const double nxt = (tt[s] + 1 < H) ? vec[lane + 32 * s + N] : 0.0;
          double r = -grad[s] - nut + (has_w ? cw[s] * iw[s] : 0.0) - (tq[s] + pg[s]) + nxt;
          X[lane + 32 * s] = r;
        }
        __syncwarp();
        solve_cols(0, 1);
        // dnu = S^{-1} (A x1 + rp): H x H Gaussian elimination by lane 0
        for (int t = 0; t < H; ++t) {
          double a = 0.0;
#pragma unroll
          for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s] && tt[s] == t) a += X[lane + 32 * s];
          a = mv_warp_sum(a);
          if (lane == 0) S[t * (H + 1) + H] = a + rp[t];
        }
        __syncwarp();
        if (lane == 0) {
          double T[MV_MAX_H][MV_MAX_H + 1];
          for (int r = 0; r < H; ++r) for (int c = 0; c <= H; ++c) T[r][c] = S[r * (H + 1) + c];
          for (int c = 0; c < H; ++c) {
            const double piv = 1.0 / T[c][c];
            for (int r = c + 1; r < H; ++r) {
              const double f = T[r][c] * piv;
              for (int c2 = c; c2 <= H; ++c2) T[r][c2] -= f * T[c][c2];
            }
          }
          for (int r = H - 1; r >= 0; --r) {
            double a = T[r][H];
            for (int c = r + 1; c < H; ++c) a -= T[r][c] * vec[c];
            vec[r] = a / T[r][r];
          }
        }
        __syncwarp();
        for (int t = 0; t < H; ++t) dnu[t] = vec[t];
        __syncwarp();
#pragma unroll
        for (int s = 0; s < MV_MAX_SLOTS; ++s) {
          double a = ok[s] ? X[lane + 32 * s] : 0.0;
          if (ok[s]) for (int t = 0; t < H; ++t) a = fma(-X[(size_t)(1 + t) * n + lane + 32 * s], dnu[t], a);
          dw[s] = a;
          if (ok[s]) vec[lane + 32 * s] = a;
        }
        __syncwarp();
        double rpm = 0.0, rdm = 0.0;
#pragma unroll
        for (int s = 0; s < MV_MAX_SLOTS; ++s) {
          if (!ok[s]) { dsp[s] = dsq[s] = dzw[s] = dzp[s] = dzq[s] = 0.0; continue; }
          const double dd = dw[s] - ((tt[s] > 0) ? vec[lane + 32 * s - N] : 0.0);
          dzw[s] = has_w ? (cw[s] * iw[s] - zw[s]) - zw[s] * iw[s] * dw[s] : 0.0;
          if (has_w) { rpm = fmax(rpm, -dw[s] * iw[s]); rdm = fmax(rdm, -dzw[s] / zw[s]); }
          if (has_u) {
            const double dp = zp[s] * isp[s], dq = zq[s] * isq[s];
            dsp[s] = (gu[s] - (2.0 * dq + delta) * dd) * ie[s];
            dsq[s] = (gu[s] + (2.0 * dp + delta) * dd) * ie[s];
            dzp[s] = (cp[s] * isp[s] - zp[s]) - dp * dsp[s];
            dzq[s] = (cq[s] * isq[s] - zq[s]) - dq * dsq[s];
            rpm = fmax(rpm, fmax(-dsp[s] * isp[s], -dsq[s] * isq[s]));
            rdm = fmax(rdm, fmax(-dzp[s] / zp[s], -dzq[s] / zq[s]));
          } else { dsp[s] = dsq[s] = dzp[s] = dzq[s] = 0.0; }
        }
        __syncwarp();
        rpm = mv_warp_max(rpm); rdm = mv_warp_max(rdm);
        aa = (mcount > 0.0 && rpm > 1.0) ? 1.0 / rpm : 1.0;
        ab = (mcount > 0.0 && rdm > 1.0) ? 1.0 / rdm : 1.0;
        if (phase == 0) {
          double g2 = 0.0;
#pragma unroll
          for (int s = 0; s < MV_MAX_SLOTS; ++s) {
            if (!ok[s]) continue;
            if (has_w) g2 = fma(fma(aa, dw[s], w[s]), fma(ab, dzw[s], zw[s]), g2);
            if (has_u) g2 = fma(fma(aa, dsp[s], sp[s]), fma(ab, dzp[s], zp[s]), fma(fma(aa, dsq[s], sq[s]), fma(ab, dzq[s], zq[s]), g2));
          }
          g2 = mv_warp_sum(g2);
          const double ratio = (gap > 0.0) ? fmin(1.0, fmax(g2 / gap, 0.0)) : 0.0;
          const double smu = ratio * ratio * ratio * mug;
          const double dmp = fmin(1.0, fmin(aa, ab) * (1.0 / kCorrFull));
#pragma unroll
          for (int s = 0; s < MV_MAX_SLOTS; ++s) {
            cw[s] = has_w ? fma(-dmp * dw[s], dzw[s], smu) : 0.0;
            cp[s] = has_u ? fma(-dmp * dsp[s], dzp[s], smu) : 0.0;
            cq[s] = has_u ? fma(-dmp * dsq[s], dzq[s], smu) : 0.0;
          }
        }
      }
      const double pa = (mcount > 0.0) ? fmin(1.0, opt.step_frac * aa) : 1.0;
      const double pb = (mcount > 0.0) ? fmin(1.0, opt.step_frac * ab) : 1.0;
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) {
        if (!ok[s]) continue;
        w[s] = fma(pa, dw[s], w[s]);
        if (has_w) zw[s] = fma(pb, dzw[s], zw[s]);
        if (has_u) {
          sp[s] = fma(pa, dsp[s], sp[s]); sq[s] = fma(pa, dsq[s], sq[s]);
          zp[s] = fma(pb, dzp[s], zp[s]); zq[s] = fma(pb, dzq[s], zq[s]);
        }
      }
      for (int t = 0; t < H; ++t) nu[t] = fma(pb, dnu[t], nu[t]);
    }
    if (status == ST_FAILED && isfinite(kkt[1] + kkt[2]) && kkt[0] < kLoosePres && kkt[1] < kLooseDres && kkt[2] < kLooseGap)
      status = ST_INACCURATE;
    // ---- outputs: plan (or tile(w_cur) on failure, mpc.py:179-180) and the maximised objective -------------------
    double val = CUDART_NAN;
    if (status <= ST_INACCURATE) {
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s]) vec[lane + 32 * s] = w[s];
      __syncwarp();
      double acc = 0.0;
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) {
        if (!ok[s]) continue;
        double a = 0.0;
        for (int j = 0; j < N; ++j) a = fma(Sig[(size_t)ii[s] * N + j], vec[tt[s] * N + j], a);
        const double prev = (tt[s] > 0) ? vec[lane + 32 * s - N] : w0[s];
        acc += w[s] * mu[s] - gamma * w[s] * a - lam * fabs(w[s] - prev);
      }
      val = mv_warp_sum(acc);
      __syncwarp();
    }
#pragma unroll
    for (int s = 0; s < MV_MAX_SLOTS; ++s)
      if (ok[s]) A.w_out[(size_t)p * n + lane + 32 * s] = (status <= ST_INACCURATE) ? w[s] : w0[s];
    if (lane == 0) {
      if (A.obj) A.obj[p] = val;
      if (A.kkt) { A.kkt[3 * p] = kkt[0]; A.kkt[3 * p + 1] = kkt[1]; A.kkt[3 * p + 2] = kkt[2]; }
      if (A.status) A.status[p] = status;
      if (A.iters) A.iters[p] = iters;
    }
    __syncwarp();
  }
}

int mv_supported(int H, int N) { return H >= 1 && H <= MV_MAX_H && N >= 1 && H * N <= 32 * MV_MAX_SLOTS; }

int launch_mpc_mv(const double* mu, const double* sigma, long long sigma_stride, const double* w_cur, double gamma, double lam,
                  int allow_short, int P, int H, int N, double* w_out, double* obj, double* kkt, int* status, int* iters,
                  int sm_count, cudaStream_t st) {
  if (!mv_supported(H, N)) return -2;
  MvArgs A;
  A.mu = mu; A.sigma = sigma; A.sigma_stride = sigma_stride; A.w_cur = w_cur; A.gamma = gamma; A.lam = lam;
  A.allow_short = allow_short; A.P = P; A.H = H; A.N = N; A.w_out = w_out; A.obj = obj; A.kkt = kkt; A.status = status;
  A.iters = iters; A.opt = default_ipm_options();
  const int n = H * N, ldm = n | 1;
  const size_t smem = ((size_t)n * ldm + (size_t)(H + 1) * n + n + (size_t)H * (H + 1) + 8) * sizeof(double);
  // the attribute is per device and monotone: raise it whenever this device has not seen a request this large
  static int attr_smem[PerDeviceInt::kMaxDevices] = {};
  static std::mutex attr_mu;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(attr_mu);
    if (dev < 0 || dev >= PerDeviceInt::kMaxDevices || (int)smem > attr_smem[dev]) {
      cudaError_t e = cudaFuncSetAttribute(mpc_mv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      if (dev >= 0 && dev < PerDeviceInt::kMaxDevices) attr_smem[dev] = (int)smem;
    }
  }
  int per_sm = (int)((size_t)220 * 1024 / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 16) per_sm = 16;
  int blocks = P < sm_count * per_sm ? P : sm_count * per_sm;
  if (blocks < 1) blocks = 1;
  mpc_mv_kernel<<<blocks, 32, smem, st>>>(A);
  return (int)cudaGetLastError();
}

}  // namespace kmpc
