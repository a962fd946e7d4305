// Mean-variance MPC (/root/reference/mpc.py:119-184, solve_mpc_mean_variance):
//
//   max  sum_t [ w_t . mu_t - gamma w_t' Sigma w_t ] - lam sum_t ||w_t - w_{t-1}||_1          (w_0 = current weights)
//   s.t. 1'w_t = 1,  w_t >= 0 (unless allow_short);  no turnover cap (mpc.py:139-176)
//
// Same fp64 primal-dual interior point as the log-utility solver (Mehrotra predictor-corrector, split steps,
// epigraph u >= |w_t - w_{t-1}| eliminated per asset, proximal term delta, loose acceptance), but the stage cost has a
// DENSE Hessian 2 gamma Sigma, so the reduced Newton matrix  M = blockdiag(2 gamma Sigma) + T  (T = the per-asset
// tridiagonal of mpc_lane.cuh) is factorised densely: in-place Cholesky, triangular solves for the right-hand side and
// the H budget columns, an H x H Schur complement for the budget multipliers.  Two instantiations of one kernel:
//   n = H N <= 160:  one WARP per problem, M (n x n) in the warp's slice of shared memory; lane l owns variables
//                    l, l+32, ... (v = t N + i);
//   n <= 1280:       one BLOCK of 256 threads per problem, M and the right-hand sides in a global-memory workspace that
//                    stays L2-resident (2 MB at N = 500); thread l owns variables l, l+256, ...
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
  double* work;           // block kernel: [blocks][work_stride] doubles for M and X
  long long work_stride;
  IpmOptions opt;
};

// The solver runs on one warp (NT = 32: M in the warp's shared memory, n <= 160) or on one block of NT = 256 threads
// (M in a global-memory workspace that stays L2-resident, n <= 5 * 256): same code, group-wide primitives.
template <int NT>
__device__ __forceinline__ void grp_sync() {
  if (NT == 32) __syncwarp(); else __syncthreads();
}
template <int NT>
__device__ __forceinline__ double grp_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  if (NT > 32) {
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    v = 0.0;
#pragma unroll
    for (int g = 0; g < NT / 32; ++g) v += red[g];
  }
  return v;
}
template <int NT>
__device__ __forceinline__ double grp_max(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
  if (NT > 32) {
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    v = red[0];
#pragma unroll
    for (int g = 1; g < NT / 32; ++g) v = fmax(v, red[g]);
  }
  return v;
}
template <int NT>
__device__ __forceinline__ bool grp_any(bool c) {
  if (NT == 32) return __any_sync(kFull, c) != 0;
  return __syncthreads_or(c ? 1 : 0) != 0;
}

// warp kernel, shared memory: M [n][ldm], X [(H+1)][n] (right-hand sides / solutions), vec [n] (neighbour exchange),
// S [H][H+1] (Schur system), red [8]
__host__ __device__ __forceinline__ size_t mv_smem_doubles(int n, int H) {
  const int ldm = n | 1;
  return (size_t)n * ldm + (size_t)(H + 1) * n + n + (size_t)H * (H + 1) + 8 + 8;
}

template <int NT>
__global__ void __launch_bounds__(NT, 1)
mpc_mv_kernel(MvArgs A) {
  extern __shared__ double smem[];
  const int lane = threadIdx.x;             // index inside the group (a warp or the block)
  const int H = A.H, N = A.N, n = H * N, ldm = n | 1;
  const int slots = (n + NT - 1) / NT;
  // NT = 32: everything in shared memory.  NT = 256: M and X in this block's slice of the global workspace.
  double* M = (NT == 32) ? smem : A.work + (size_t)blockIdx.x * A.work_stride;
  double* X = M + (size_t)n * ldm;          // X[c*n + v], c = 0: rhs, c = 1 + t: budget column of stage t
  double* vec = (NT == 32) ? X + (size_t)(H + 1) * n : smem;
  double* S = vec + n;
  double* red = S + (size_t)H * (H + 1) + 8;
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
      const int v = lane + NT * s;
      ok[s] = (s < slots) && (v < n);
      tt[s] = ok[s] ? v / N : 0; ii[s] = ok[s] ? v - tt[s] * N : 0;
      mu[s] = ok[s] ? A.mu[(size_t)p * n + v] : 0.0;
      w0[s] = ok[s] ? A.w_cur[(size_t)p * N + ii[s]] : 0.0;
      if (ok[s] && !(isfinite(mu[s]) && isfinite(w0[s]))) bad = true;
    }
    int status = ST_FAILED, iters = 0;
    double kkt[3] = {CUDART_NAN, CUDART_NAN, CUDART_NAN};
    if (grp_any<NT>(bad)) status = ST_NONFINITE;
    // ---- initial point (oracle _initial_point with tau = 0; dual-feasible start as in mpc_lane.cuh) -------------
    double delta = opt.delta;
    if (status != ST_NONFINITE) {
      double sb = 0.0;
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s] && tt[s] == 0) sb += A.allow_short ? w0[s] : fmax(w0[s], 0.0);
      sb = grp_sum<NT>(sb, red);
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
      for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s]) vec[lane + NT * s] = w[s];
      grp_sync<NT>();
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
        nu[t] = grp_max<NT>(mx, red) + (has_w ? opt.dual_init : 0.0);
      }
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) {
        double nut = 0.0;
        for (int t = 0; t < H; ++t) if (t == tt[s]) nut = nu[t];
        zw[s] = (has_w && ok[s]) ? grad[s] + nut : 0.0;
      }
      grp_sync<NT>();
    }
    const double mcount = (has_w ? (double)n : 0.0) + (has_u ? 2.0 * n : 0.0);
    // ---- iterations ---------------------------------------------------------------------------------------------
    for (int it = 1; status == ST_FAILED && it <= opt.max_iter + 1; ++it) {
      iters = it;
      // residuals: Sigma w (dense), neighbours through vec
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s]) vec[lane + NT * s] = w[s];
      grp_sync<NT>();
      double grad[MV_MAX_SLOTS], y[MV_MAX_SLOTS];
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) {
        double a = 0.0;
        if (ok[s]) for (int j = 0; j < N; ++j) a = fma(Sig[(size_t)ii[s] * N + j], vec[tt[s] * N + j], a);
        grad[s] = fma(2.0 * gamma, a, -mu[s]);
        y[s] = zp[s] - zq[s];
      }
      grp_sync<NT>();
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s]) vec[lane + NT * s] = y[s];
      grp_sync<NT>();
      double dres = 0.0, gap = 0.0, pres = 0.0;
      double rp[MV_MAX_H];
      for (int t = 0; t < H; ++t) {
        double sw = 0.0;
#pragma unroll
        for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s] && tt[s] == t) sw += w[s];
        rp[t] = grp_sum<NT>(sw, red) - 1.0;
        pres = fmax(pres, fabs(rp[t]));
      }
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) {
        if (!ok[s]) continue;
        const double yn = (tt[s] + 1 < H) ? vec[lane + NT * s + N] : 0.0;
        double nut = 0.0;
        for (int t = 0; t < H; ++t) if (t == tt[s]) nut = nu[t];
        dres = fmax(dres, fabs(grad[s] - (has_w ? zw[s] : 0.0) + y[s] - yn + nut));
        if (has_u) dres = fmax(dres, fabs(lam - zp[s] - zq[s]));
        if (has_w) gap = fma(w[s], zw[s], gap);
        if (has_u) gap = fma(sp[s], zp[s], fma(sq[s], zq[s], gap));
      }
      grp_sync<NT>();
      dres = grp_max<NT>(dres, red); gap = grp_sum<NT>(gap, red);
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
        if (ok[s]) vec[lane + NT * s] = ee[s];
      }
      grp_sync<NT>();
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) {
        if (!ok[s]) continue;
        const int v = lane + NT * s;
        double* row = M + (size_t)v * ldm;
        for (int c = 0; c <= v; ++c) row[c] = 0.0;
        const int base = tt[s] * N;
        for (int j = 0; j <= ii[s]; ++j) row[base + j] = 2.0 * gamma * Sig[(size_t)ii[s] * N + j];
        const double en = (tt[s] + 1 < H) ? vec[v + N] : 0.0;            // edge to the next stage
        row[v] += ad[s] + ee[s] + en;
        if (tt[s] > 0) row[v - N] = -ee[s];                                // edge to the previous stage
      }
      grp_sync<NT>();
      // ---- in-place Cholesky (lower) ---------------------------------------------------------------------------------
      bool pd = true;
      for (int j = 0; j < n; ++j) {
        const double d = M[(size_t)j * ldm + j];
        if (!(d > 0.0)) { pd = false; break; }
        const double inv = rsqrt(d);
        for (int i = j + lane; i < n; i += NT) M[(size_t)i * ldm + j] = (i == j) ? d * inv : M[(size_t)i * ldm + j] * inv;
        grp_sync<NT>();
        for (int i = j + 1 + lane; i < n; i += NT) {
          const double lij = M[(size_t)i * ldm + j];
          double* ri = M + (size_t)i * ldm;
          for (int k = j + 1; k <= i; ++k) ri[k] = fma(-lij, M[(size_t)k * ldm + j], ri[k]);
        }
        grp_sync<NT>();
      }
      if (!pd) break;
      // budget columns X[1+t] = M^{-1} A_t' (once per factorisation)
      auto solve_cols = [&](int c0, int c1) {          // in-place L L' solves of columns c0..c1-1 of X
        for (int j = 0; j < n; ++j) {
          const double invd = 1.0 / M[(size_t)j * ldm + j];
          for (int c = c0; c < c1; ++c) {
            double* x = X + (size_t)c * n;
            const double yj = x[j] * invd;
            grp_sync<NT>();
            if (lane == 0) x[j] = yj;
            for (int i = j + 1 + lane; i < n; i += NT) x[i] = fma(-M[(size_t)i * ldm + j], yj, x[i]);
          }
          grp_sync<NT>();
        }
        for (int j = n - 1; j >= 0; --j) {
          const double invd = 1.0 / M[(size_t)j * ldm + j];
          for (int c = c0; c < c1; ++c) {
            double* x = X + (size_t)c * n;
            const double xj = x[j] * invd;
            grp_sync<NT>();
            if (lane == 0) x[j] = xj;
            for (int i = lane; i < j; i += NT) x[i] = fma(-M[(size_t)j * ldm + i], xj, x[i]);
          }
          grp_sync<NT>();
        }
      };
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s)
        if (ok[s]) for (int t = 0; t < H; ++t) X[(size_t)(1 + t) * n + lane + NT * s] = (t == tt[s]) ? 1.0 : 0.0;
      grp_sync<NT>();
      solve_cols(1, 1 + H);
      // Schur matrix  S[t][t2] = A_t X2[:, t2] = sum over stage-t variables of column t2
      for (int t = 0; t < H; ++t)
        for (int t2 = 0; t2 < H; ++t2) {
          double a = 0.0;
#pragma unroll
          for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s] && tt[s] == t) a += X[(size_t)(1 + t2) * n + lane + NT * s];
          a = grp_sum<NT>(a, red);
          if (lane == 0) S[t * (H + 1) + t2] = a;
        }
      grp_sync<NT>();
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
          if (ok[s]) vec[lane + NT * s] = tq[s] + pg[s];          // both enter as  - x_t + x_{t+1}
        }
        grp_sync<NT>();
#pragma unroll
        for (int s = 0; s < MV_MAX_SLOTS; ++s) {
          if (!ok[s]) continue;
          double nut = 0.0;
          for (int t = 0; t < H; ++t) if (t == tt[s]) nut = nu[t];
          const double nxt = (tt[s] + 1 < H) ? vec[lane + NT * s + N] : 0.0;
          double r = -grad[s] - nut + (has_w ? cw[s] * iw[s] : 0.0) - (tq[s] + pg[s]) + nxt;
          X[lane + NT * s] = r;
        }
        grp_sync<NT>();
        solve_cols(0, 1);
        // dnu = S^{-1} (A x1 + rp): H x H Gaussian elimination by lane 0
        for (int t = 0; t < H; ++t) {
          double a = 0.0;
#pragma unroll
          for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s] && tt[s] == t) a += X[lane + NT * s];
          a = grp_sum<NT>(a, red);
          if (lane == 0) S[t * (H + 1) + H] = a + rp[t];
        }
        grp_sync<NT>();
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
        grp_sync<NT>();
        for (int t = 0; t < H; ++t) dnu[t] = vec[t];
        grp_sync<NT>();
#pragma unroll
        for (int s = 0; s < MV_MAX_SLOTS; ++s) {
          double a = ok[s] ? X[lane + NT * s] : 0.0;
          if (ok[s]) for (int t = 0; t < H; ++t) a = fma(-X[(size_t)(1 + t) * n + lane + NT * s], dnu[t], a);
          dw[s] = a;
          if (ok[s]) vec[lane + NT * s] = a;
        }
        grp_sync<NT>();
        double rpm = 0.0, rdm = 0.0;
#pragma unroll
        for (int s = 0; s < MV_MAX_SLOTS; ++s) {
          if (!ok[s]) { dsp[s] = dsq[s] = dzw[s] = dzp[s] = dzq[s] = 0.0; continue; }
          const double dd = dw[s] - ((tt[s] > 0) ? vec[lane + NT * s - N] : 0.0);
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
        grp_sync<NT>();
        rpm = grp_max<NT>(rpm, red); rdm = grp_max<NT>(rdm, red);
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
          g2 = grp_sum<NT>(g2, red);
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
      for (int s = 0; s < MV_MAX_SLOTS; ++s) if (ok[s]) vec[lane + NT * s] = w[s];
      grp_sync<NT>();
      double acc = 0.0;
#pragma unroll
      for (int s = 0; s < MV_MAX_SLOTS; ++s) {
        if (!ok[s]) continue;
        double a = 0.0;
        for (int j = 0; j < N; ++j) a = fma(Sig[(size_t)ii[s] * N + j], vec[tt[s] * N + j], a);
        const double prev = (tt[s] > 0) ? vec[lane + NT * s - N] : w0[s];
        acc += w[s] * mu[s] - gamma * w[s] * a - lam * fabs(w[s] - prev);
      }
      val = grp_sum<NT>(acc, red);
      grp_sync<NT>();
    }
#pragma unroll
    for (int s = 0; s < MV_MAX_SLOTS; ++s)
      if (ok[s]) A.w_out[(size_t)p * n + lane + NT * s] = (status <= ST_INACCURATE) ? w[s] : w0[s];
    if (lane == 0) {
      if (A.obj) A.obj[p] = val;
      if (A.kkt) { A.kkt[3 * p] = kkt[0]; A.kkt[3 * p + 1] = kkt[1]; A.kkt[3 * p + 2] = kkt[2]; }
      if (A.status) A.status[p] = status;
      if (A.iters) A.iters[p] = iters;
    }
    grp_sync<NT>();
  }
}

constexpr int MV_BLOCK_THREADS = 256;
int mv_supported(int H, int N) { return H >= 1 && H <= MV_MAX_H && N >= 1 && H * N <= MV_BLOCK_THREADS * MV_MAX_SLOTS; }
// doubles of global workspace the block kernel needs per block (0: the warp kernel takes this shape)
long long mv_work_doubles(int H, int N) {
  const long long n = (long long)H * N, ldm = n | 1;
  return (n <= 32 * MV_MAX_SLOTS) ? 0 : n * ldm + (long long)(H + 1) * n;
}
int mv_blocks(int P, int H, int N, int sm_count) {
  if (mv_work_doubles(H, N) == 0) return 0;
  return P < sm_count ? P : sm_count;
}

int launch_mpc_mv(const double* mu, const double* sigma, long long sigma_stride, const double* w_cur, double gamma, double lam,
                  int allow_short, int P, int H, int N, double* w_out, double* obj, double* kkt, int* status, int* iters,
                  double* work, int sm_count, cudaStream_t st) {
  if (!mv_supported(H, N)) return -2;
  MvArgs A;
  A.mu = mu; A.sigma = sigma; A.sigma_stride = sigma_stride; A.w_cur = w_cur; A.gamma = gamma; A.lam = lam;
  A.allow_short = allow_short; A.P = P; A.H = H; A.N = N; A.w_out = w_out; A.obj = obj; A.kkt = kkt; A.status = status;
  A.iters = iters; A.opt = default_ipm_options(); A.work = work; A.work_stride = mv_work_doubles(H, N);
  const int n = H * N;
  if (A.work_stride > 0) {
    // large problems: one block of 256 threads per problem, M and X in the caller's global workspace (L2-resident)
    if (!work) return -2;
    const size_t smem = ((size_t)n + (size_t)H * (H + 1) + 8 + 8) * sizeof(double);
    mpc_mv_kernel<MV_BLOCK_THREADS><<<mv_blocks(P, H, N, sm_count), MV_BLOCK_THREADS, smem, st>>>(A);
    return (int)cudaGetLastError();
  }
  const size_t smem = mv_smem_doubles(n, H) * sizeof(double);
  // the attribute is per device and monotone: raise it whenever this device has not seen a request this large
  static int attr_smem[PerDeviceInt::kMaxDevices] = {};
  static std::mutex attr_mu;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(attr_mu);
    if (dev < 0 || dev >= PerDeviceInt::kMaxDevices || (int)smem > attr_smem[dev]) {
      cudaError_t e = cudaFuncSetAttribute(mpc_mv_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
      if (dev >= 0 && dev < PerDeviceInt::kMaxDevices) attr_smem[dev] = (int)smem;
    }
  }
  int per_sm = (int)((size_t)220 * 1024 / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 16) per_sm = 16;
  int blocks = P < sm_count * per_sm ? P : sm_count * per_sm;
  if (blocks < 1) blocks = 1;
  mpc_mv_kernel<32><<<blocks, 32, smem, st>>>(A);
  return (int)cudaGetLastError();
}

}  // namespace kmpc
