// Warp-per-problem fp64 primal-dual interior-point solver for the MPC program of
// /root/reference/mpc.py:27-117 (solve_mpc_log_utility):
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
//     whose SPD Schur complement K is assembled with warp reductions and factorised by the warp.
// The numpy twin of this file, iteration for iteration, is oracle/mpc_oracle.py::solve_structured.
//
// Lane l owns assets l, l+32, ... (APT per lane); H is a compile-time constant so that every
// per-stage array lives in registers (or compiler-managed local memory when it does not fit).
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
};

__host__ __device__ inline IpmOptions default_ipm_options() {
  IpmOptions o;
  o.tol = 1e-10; o.tol_dual = 1e-8; o.delta = 1e-5; o.step_frac = 0.995; o.mu0 = 1e-3; o.dual_init = 3e-3;
  o.max_iter = 50;
  return o;
}

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(kFull, v, m); }
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(kFull, v, src); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
  return v;
}
// min / max of non-negative values through the integer REDUX unit (fp32 precision, rounded safely)
__device__ __forceinline__ double warp_min_pos(double v) {
  float f = __double2float_rd(v);
  unsigned u = __reduce_min_sync(kFull, __float_as_uint(f));
  return (double)__uint_as_float(u);
}
__device__ __forceinline__ double warp_max_pos(double v) {
  float f = __double2float_ru(v);
  unsigned u = __reduce_max_sync(kFull, __float_as_uint(f));
  return (double)__uint_as_float(u);
}

// Reduce NV (<= 32) per-lane values across the warp; the total of entry e lands in lane e.
// 31 double shuffles for 32 entries instead of 160.
template <int NV>
__device__ __forceinline__ double warp_transpose_reduce(double (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    if (NV > o) {
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int i = 0; i < o; ++i) {
        if (i + o < NV || i < NV) {
          double lo = (i < NV) ? v[i] : 0.0;
          double hi = (i + o < NV) ? v[i + o] : 0.0;
          double send = up ? lo : hi;
          double keep = up ? hi : lo;
          v[i] = keep + shfl_xor_d(send, o);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] += shfl_xor_d(v[i], o);
    }
  }
  return v[0];
}

template <int H, int APT>
struct WarpIpm {
  static constexpr int NB = 3 * H;          // border size (R-rows, budget rows, cap rows)
  static constexpr int SMEM_DOUBLES = NB * NB + NB;
  static_assert(NB <= 32, "border must fit one entry per lane");

  // ---- per-lane state -------------------------------------------------------------------------
  double R[APT][H], w[APT][H], sp[APT][H], sq[APT][H], zw[APT][H], zp[APT][H], zq[APT][H];
  double nu[H], sc[H], zc[H];               // warp-uniform
  bool valid[APT];
  // ---- factorisation of the current iterate -----------------------------------------------------
  double qL[APT][H], tL[APT][H], qR[APT][H], tR[APT][H], gjj[APT][H], Vd[APT][H], fL[APT][H], fR[APT][H];
  double phi[APT][H], iE[APT][H], Dp[APT][H], Dq[APT][H], Dw0[APT][H];
  double* Ksm;                              // [NB*NB] lower Cholesky factor (row-major), then [NB] 1/diag
  int lane;
  bool has_w, has_u, has_c;
  double lam, tau, delta;
  int nb;                                   // active border size: 2H or 3H

  // M0^{-1} (g_w, g_u): dw, dd through the Green's functions; pg = phi * g_u (dipole strengths)
  __device__ __forceinline__ void m0_apply(int a, const double (&gwv)[H], const double (&pg)[H],
                                           double (&dw)[H], double (&dd)[H]) const {
#pragma unroll
    for (int k = 0; k < H; ++k) { dw[k] = 0.0; dd[k] = 0.0; }
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const double g = gwv[j];
      double acc = 0.0;                     // sum_k D[k,j] * pg[k]  (potential of node j from the dipoles)
      double v = gjj[a][j];
      dw[j] += v * g;
#pragma unroll
      for (int l = j + 1; l < H; ++l) {     // propagate right
        const double Dlj = -v * qR[a][l];
        v *= tR[a][l];
        dw[l] += v * g;
        dd[l] += Dlj * g;
        acc += Dlj * pg[l];
      }
      v = gjj[a][j];
#pragma unroll
      for (int l = j; l >= 0; --l) {        // propagate left
        const double Dlj = v * qL[a][l];
        dd[l] += Dlj * g;
        acc += Dlj * pg[l];
        v *= tL[a][l];
        if (l >= 1) dw[l - 1] += v * g;
      }
      dw[j] -= acc;
    }
    if (has_u) {
#pragma unroll
      for (int k = 0; k < H; ++k) {         // dipole across edge k
        const double p = pg[k];
        const double V = Vd[a][k];
        dd[k] -= V * p;
        double v = V * fL[a][k];
#pragma unroll
        for (int l = k + 1; l < H; ++l) {
          dd[l] += v * qR[a][l] * p;        // DD[l,k] = -v*qR
          v *= tR[a][l];
        }
        v = -V * fR[a][k];
#pragma unroll
        for (int l = k - 1; l >= 0; --l) {
          dd[l] -= v * qL[a][l] * p;        // DD[l,k] = v*qL
          v *= tL[a][l];
        }
      }
    }
  }

  // Build the Green's-function factors of every owned asset and the border matrix K; Cholesky.
  // Returns false on a non-positive pivot.
  __device__ bool factorize(const double (&rho)[H]) {
    double c[32];
    // K entry numbering (lower triangle incl. diagonal, row-major over the active nb x nb border)
    // is processed in batches of 32 partial sums.
    const int ntri = nb * (nb + 1) / 2;
    // per-asset Green factors
#pragma unroll
    for (int a = 0; a < APT; ++a) {
      double e[H], hLv[H], hRv[H], ad[H];
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const double dw0 = (has_w && valid[a]) ? zw[a][k] / w[a][k] : 0.0;
        Dw0[a][k] = dw0;
        ad[k] = dw0 + delta;
        if (has_u) {
          const double dp = zp[a][k] / sp[a][k], dq = zq[a][k] / sq[a][k];
          const double E = dp + dq + delta;
          const double ie = 1.0 / E;
          Dp[a][k] = dp; Dq[a][k] = dq; iE[a][k] = ie;
          phi[a][k] = (dq - dp) * ie;
          e[k] = (4.0 * dp * dq + 2.0 * delta * (dp + dq) + delta * delta) * ie;
        } else {
          Dp[a][k] = 0.0; Dq[a][k] = 0.0; iE[a][k] = 1.0; phi[a][k] = 0.0; e[k] = 0.0;
        }
      }
      // left sweep: hL[k] = conductance to ground seen at node k leftwards incl. ad[k]
#pragma unroll
      for (int l = 0; l < H; ++l) {
        if (l == 0) { qL[a][0] = 1.0; tL[a][0] = 0.0; }
        else {
          const double inv = 1.0 / (e[l] + hLv[l - 1]);
          qL[a][l] = hLv[l - 1] * inv; tL[a][l] = e[l] * inv;
        }
        hLv[l] = ad[l] + e[l] * qL[a][l];
      }
      // right sweep: hR[k] incl. ad[k]; qR/tR indexed by the edge entering node l from the left
      hRv[H - 1] = ad[H - 1];
#pragma unroll
      for (int l = H - 1; l >= 1; --l) {
        const double inv = 1.0 / (e[l] + hRv[l]);
        qR[a][l] = hRv[l] * inv; tR[a][l] = e[l] * inv;
        hRv[l - 1] = ad[l - 1] + e[l] * qR[a][l];
      }
      qR[a][0] = 0.0; tR[a][0] = 0.0;
#pragma unroll
      for (int j = 0; j < H; ++j) {
        const double gR = (j + 1 < H) ? e[j + 1] * qR[a][j + 1] : 0.0;
        gjj[a][j] = 1.0 / (hLv[j] + gR);
        double fl = 1.0, fr = 0.0;
        if (j > 0) {
          const double inv = 1.0 / (hLv[j - 1] + hRv[j]);
          fl = hLv[j - 1] * inv; fr = hRv[j] * inv;
        }
        fL[a][j] = fl; fR[a][j] = fr;
        Vd[a][j] = 1.0 / (e[j] + hRv[j] * fl);
      }
    }
    // K assembly: entries in batches of 32; entry (r, cidx) with r >= cidx (lower triangle).
    // Border rows: [0,H) = Rt_k, [H,2H) = 1t_k, [2H,3H) = et_k.
    int r0 = 0, c0 = 0;                                // (row, col) of entry `base`
    for (int base = 0; base < ntri; base += 32) {
      int r = r0, cc = c0;
      int myr = 0, myc = 0;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        double s = 0.0;
        if (base + i < ntri) {
          const int rt = r / H, rl = r - rt * H;       // type and stage of the row / column
          const int ct = cc / H, cl = cc - ct * H;
#pragma unroll
          for (int a = 0; a < APT; ++a)
            if (valid[a]) s += k_entry(a, rt, rl, ct, cl);
        }
        c[i] = s;
        if (i == lane) { myr = r; myc = cc; }
        if (++cc > r) { ++r; cc = 0; }
      }
      r0 = r; c0 = cc;
      const double tot = warp_transpose_reduce<32>(c, lane);
      if (base + lane < ntri) Ksm[myr * NB + myc] = tot;
    }
    __syncwarp();
    if (lane < H) {
      double rk = 0.0, sk = 0.0;
#pragma unroll
      for (int k = 0; k < H; ++k) if (lane == k) { rk = rho[k]; sk = has_c ? sc[k] / zc[k] : 0.0; }
      Ksm[lane * NB + lane] += rk * rk;                                     // 1/beta_k
      if (has_c) Ksm[(2 * H + lane) * NB + 2 * H + lane] += sk;
    }
    __syncwarp();
    // warp Cholesky (lower, in place); lane i owns row i
    bool ok = true;
    for (int j = 0; j < nb; ++j) {
      const double djj = Ksm[j * NB + j];
      if (!(djj > 0.0)) { ok = false; break; }
      const double inv = 1.0 / sqrt(djj);
      if (lane > j && lane < nb) Ksm[lane * NB + j] *= inv;
      if (lane == j) { Ksm[j * NB + j] = djj * inv; Ksm[NB * NB + j] = inv; }
      __syncwarp();
      if (lane > j && lane < nb) {
        const double lij = Ksm[lane * NB + j];
        for (int k2 = j + 1; k2 <= lane; ++k2) Ksm[lane * NB + k2] -= lij * Ksm[k2 * NB + j];
      }
      __syncwarp();
    }
    return ok;
  }

  // Green's function value needed by one K entry for asset slot a.
  // G[l,j]: node l potential for unit injection at j;  D[l,j]: drop across edge l;  DD[l,k].
  __device__ __forceinline__ double green_G(int a, int l, int j) const {
    double v = gjj[a][j];
    if (l > j) { for (int m = j + 1; m <= l; ++m) v *= tR[a][m]; }
    else { for (int m = j; m > l; --m) v *= tL[a][m]; }
    return v;
  }
  __device__ __forceinline__ double green_D(int a, int l, int j) const {
    if (l <= j) return green_G(a, l, j) * qL[a][l];
    return -green_G(a, l - 1, j) * qR[a][l];
  }
  __device__ __forceinline__ double green_DD(int a, int l, int k) const {
    const double V = Vd[a][k];
    if (l == k) return V;
    if (l > k) {
      double v = V * fL[a][k];
      for (int m = k + 1; m < l; ++m) v *= tR[a][m];
      return -v * qR[a][l];
    }
    double v = -V * fR[a][k];
    for (int m = k - 1; m > l; --m) v *= tL[a][m];
    return v * qL[a][l];
  }
  __device__ __forceinline__ double k_entry(int a, int rt, int rl, int ct, int cl) const {
    // row type rt in {0:R,1:one,2:cap}, col type ct <= rt ordering not guaranteed; handle all pairs
    if (rt < 2 && ct < 2) {
      const double g = green_G(a, rl, cl);
      return g * (rt == 0 ? R[a][rl] : 1.0) * (ct == 0 ? R[a][cl] : 1.0);
    }
    if (rt == 2 && ct < 2) {   // S[et_l, Rt_j or 1t_j] = sum -phi_l D[l,j] * (R_j or 1)
      return -phi[a][rl] * green_D(a, rl, cl) * (ct == 0 ? R[a][cl] : 1.0);
    }
    // rt == 2 && ct == 2
    double v = phi[a][rl] * phi[a][cl] * green_DD(a, rl, cl);
    if (rl == cl) v += iE[a][rl];
    return v;
  }

  // y <- K^{-1} t ; lane i holds t_i on entry and y_i on exit (i < nb)
  __device__ __forceinline__ double k_solve(double t) const {
    for (int j = 0; j < nb; ++j) {                 // forward: L y = t
      const double yj = shfl_d(t, j) * Ksm[NB * NB + j];
      if (lane == j) t = yj;
      if (lane > j && lane < nb) t -= Ksm[lane * NB + j] * yj;
    }
    for (int j = nb - 1; j >= 0; --j) {            // backward: L' x = y
      const double xj = shfl_d(t, j) * Ksm[NB * NB + j];
      if (lane == j) t = xj;
      if (lane < j) t -= Ksm[j * NB + lane] * xj;
    }
    return t;
  }

  struct Dir {
    double dw[APT][H], dsp[APT][H], dsq[APT][H], dzw[APT][H], dzp[APT][H], dzq[APT][H];
    double dnu[H], dsc[H], dzc[H];
  };

  // One Newton solve with complementarity targets c* (c = sigma*mu - corrector products).
  __device__ void newton(const double (&gw)[APT][H], const double (&rp)[H],
                         const double (&cw)[APT][H], const double (&cp)[APT][H], const double (&cq)[APT][H],
                         const double (&cc)[H], Dir& d) const {
    double g_w[APT][H], g_u[APT][H];
    double tv[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) tv[i] = 0.0;
    // right-hand side  rhs_x = -grad f - A' nu - G'(c/s)
#pragma unroll
    for (int a = 0; a < APT; ++a) {
      double tq[H];
#pragma unroll
      for (int k = 0; k < H; ++k) {
        double gwk = -gw[a][k] - nu[k];
        if (has_w) gwk += cw[a][k] / w[a][k];
        double guk = 0.0;
        tq[k] = 0.0;
        if (has_u) {
          const double a1 = cp[a][k] / sp[a][k], a2 = cq[a][k] / sq[a][k];
          tq[k] = a1 - a2;
          guk = -lam + a1 + a2;
          if (has_c) guk -= cc[k] / sc[k];
        }
        g_w[a][k] = gwk; g_u[a][k] = guk;
      }
#pragma unroll
      for (int k = 0; k < H; ++k) {
        g_w[a][k] -= tq[k];
        if (k + 1 < H) g_w[a][k] += tq[k + 1];
      }
      if (!valid[a]) {
#pragma unroll
        for (int k = 0; k < H; ++k) { g_w[a][k] = 0.0; g_u[a][k] = 0.0; }
      }
    }
    // first pass: t = V' M0^{-1} g
#pragma unroll
    for (int a = 0; a < APT; ++a) {
      if (!valid[a]) continue;
      double pg[H], dw0[H], dd0[H];
#pragma unroll
      for (int k = 0; k < H; ++k) pg[k] = phi[a][k] * g_u[a][k];
      m0_apply(a, g_w[a], pg, dw0, dd0);
#pragma unroll
      for (int k = 0; k < H; ++k) {
        tv[k] += R[a][k] * dw0[k];
        tv[H + k] += dw0[k];
        if (has_c) tv[2 * H + k] += g_u[a][k] * iE[a][k] - phi[a][k] * dd0[k];   // du0
      }
    }
    double t = warp_transpose_reduce<NB>(tv, lane);
#pragma unroll
    for (int k = 0; k < H; ++k) if (lane == H + k) t += rp[k];   // t[H+k] = sum dw0 - q, q = -rp
    const double y = k_solve(t);
    // second pass: dx = M0^{-1}(g - V y)
    double yR[H], yN[H], yC[H];
#pragma unroll
    for (int k = 0; k < H; ++k) {
      yR[k] = shfl_d(y, k);
      yN[k] = shfl_d(y, H + k);
      yC[k] = has_c ? shfl_d(y, 2 * H + k) : 0.0;
      d.dnu[k] = yN[k];
      if (has_c) {
        d.dsc[k] = -yC[k] * sc[k] / zc[k];
        d.dzc[k] = (cc[k] / sc[k] - zc[k]) + yC[k];
      } else { d.dsc[k] = 0.0; d.dzc[k] = 0.0; }
    }
#pragma unroll
    for (int a = 0; a < APT; ++a) {
      double gw2[H], geff[H], pg[H], dw[H], dd[H];
#pragma unroll
      for (int k = 0; k < H; ++k) {
        gw2[k] = g_w[a][k] - yR[k] * R[a][k] - yN[k];
        geff[k] = g_u[a][k] - yC[k];
        pg[k] = phi[a][k] * geff[k];
      }
      if (valid[a]) m0_apply(a, gw2, pg, dw, dd);
#pragma unroll
      for (int k = 0; k < H; ++k) {
        if (!valid[a]) {
          d.dw[a][k] = 0.0; d.dsp[a][k] = 0.0; d.dsq[a][k] = 0.0; d.dzw[a][k] = 0.0; d.dzp[a][k] = 0.0; d.dzq[a][k] = 0.0;
          continue;
        }
        d.dw[a][k] = dw[k];
        d.dzw[a][k] = has_w ? (cw[a][k] / w[a][k] - zw[a][k]) - Dw0[a][k] * dw[k] : 0.0;
        if (has_u) {
          const double dsp_ = (geff[k] - (2.0 * Dq[a][k] + delta) * dd[k]) * iE[a][k];
          const double dsq_ = (geff[k] + (2.0 * Dp[a][k] + delta) * dd[k]) * iE[a][k];
          d.dsp[a][k] = dsp_; d.dsq[a][k] = dsq_;
          d.dzp[a][k] = (cp[a][k] / sp[a][k] - zp[a][k]) - Dp[a][k] * dsp_;
          d.dzq[a][k] = (cq[a][k] / sq[a][k] - zq[a][k]) - Dq[a][k] * dsq_;
        } else {
          d.dsp[a][k] = 0.0; d.dsq[a][k] = 0.0; d.dzp[a][k] = 0.0; d.dzq[a][k] = 0.0;
        }
      }
    }
  }

  // largest steps keeping slacks (ap) and duals (ad) positive
  __device__ void max_step(const Dir& d, const double (&rho)[H], bool allow_short, double& ap, double& ad) const {
    double p = 1.0, q = 1.0;
    auto lim = [](double v, double dv, double a0) { return (dv < 0.0) ? fmin(a0, -v / dv) : a0; };
#pragma unroll
    for (int a = 0; a < APT; ++a) {
      if (!valid[a]) continue;
#pragma unroll
      for (int k = 0; k < H; ++k) {
        if (has_w) { p = lim(w[a][k], d.dw[a][k], p); q = lim(zw[a][k], d.dzw[a][k], q); }
        if (has_u) {
          p = lim(sp[a][k], d.dsp[a][k], p); p = lim(sq[a][k], d.dsq[a][k], p);
          q = lim(zp[a][k], d.dzp[a][k], q); q = lim(zq[a][k], d.dzq[a][k], q);
        }
      }
    }
    if (allow_short) {       // keep the log argument positive
      double dr[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) dr[i] = 0.0;
#pragma unroll
      for (int a = 0; a < APT; ++a)
        if (valid[a]) {
#pragma unroll
          for (int k = 0; k < H; ++k) dr[k] += d.dw[a][k] * R[a][k];
        }
      const double tot = warp_transpose_reduce<H>(dr, lane);
#pragma unroll
      for (int k = 0; k < H; ++k) p = lim(rho[k], shfl_d(tot, k), p);
    }
    if (has_c) {
#pragma unroll
      for (int k = 0; k < H; ++k) { p = lim(sc[k], d.dsc[k], p); q = lim(zc[k], d.dzc[k], q); }
    }
    // exact fp64 min over the warp (the fp32 REDUX shortcut would perturb the iterates w.r.t. the oracle)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { p = fmin(p, shfl_xor_d(p, o)); q = fmin(q, shfl_xor_d(q, o)); }
    ap = p; ad = q;
  }

  // Solve.  On entry R, valid, lane, Ksm are set.  w0[a] = current weights of the owned assets.
  // Returns status; w[][] holds the plan (or tile(w0) on failure), kkt = (pres, dres, gap).
  __device__ int solve(const double (&w0)[APT], int N, double lam_, double tau_, bool allow_short,
                       const IpmOptions& opt, int& iters, double (&kkt)[3]) {
    lam = lam_; tau = tau_; delta = opt.delta;
    has_u = (lam > 0.0) || (tau > 0.0);
    has_c = has_u && (tau > 0.0);
    has_w = !allow_short;
    nb = has_c ? 3 * H : 2 * H;
    iters = 0;
    kkt[0] = kkt[1] = kkt[2] = CUDART_NAN;
    // ---- input screening -------------------------------------------------------------------------
    int bad = 0;
#pragma unroll
    for (int a = 0; a < APT; ++a)
      if (valid[a]) {
        if (!isfinite(w0[a])) bad = 1;
#pragma unroll
        for (int k = 0; k < H; ++k) if (!(isfinite(R[a][k]) && R[a][k] > 0.0)) bad = 1;
      }
    if (__any_sync(kFull, bad)) { hold(w0); return ST_NONFINITE; }
    // ---- initial point (oracle/mpc_oracle.py::_initial_point) ----------------------------------------
    double base[APT], sb = 0.0;
#pragma unroll
    for (int a = 0; a < APT; ++a) {
      base[a] = valid[a] ? (allow_short ? w0[a] : fmax(w0[a], 0.0)) : 0.0;
      sb += base[a];
    }
    sb = warp_sum(sb);
    const double invN = 1.0 / (double)N;
    const double eps = (tau <= 0.0) ? 0.1 : fmin(0.1, tau / 8.0);
    double absd0 = 0.0;
#pragma unroll
    for (int a = 0; a < APT; ++a) {
      const double b = (sb > 0.0) ? base[a] / sb : invN;
      const double w1 = valid[a] ? (1.0 - eps) * b + eps * invN : 1.0;
#pragma unroll
      for (int k = 0; k < H; ++k) w[a][k] = w1;
      if (valid[a]) absd0 += fabs(w1 - w0[a]);
    }
    absd0 = warp_sum(absd0);
    if (has_u) {
      double dl0, dlk;
      if (tau > 0.0) {
        const double room0 = tau - absd0;
        if (!(room0 > 0.0)) { hold(w0); kkt[0] = kkt[1] = kkt[2] = CUDART_INF; return ST_FAILED; }
        dl0 = room0 / (2.0 * N); dlk = tau / (2.0 * N);
      } else { dl0 = dlk = 0.05 * invN; }
      double su0 = 0.0;
#pragma unroll
      for (int a = 0; a < APT; ++a) {
        const double d0 = valid[a] ? w[a][0] - w0[a] : 0.0;
        const double u0 = fabs(d0) + dl0;
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const double dk = (k == 0) ? d0 : 0.0;
          const double uk = (k == 0) ? u0 : dlk;
          sp[a][k] = uk - dk; sq[a][k] = uk + dk;
        }
        if (valid[a]) su0 += u0;
      }
      su0 = warp_sum(su0);
#pragma unroll
      for (int k = 0; k < H; ++k) sc[k] = has_c ? (tau - ((k == 0) ? su0 : dlk * N)) : 1.0;
    } else {
#pragma unroll
      for (int a = 0; a < APT; ++a)
#pragma unroll
        for (int k = 0; k < H; ++k) { sp[a][k] = 1.0; sq[a][k] = 1.0; }
#pragma unroll
      for (int k = 0; k < H; ++k) sc[k] = 1.0;
    }
    double rho[H];
    {
      double rs[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) rs[i] = 0.0;
#pragma unroll
      for (int a = 0; a < APT; ++a)
        if (valid[a]) {
#pragma unroll
          for (int k = 0; k < H; ++k) rs[k] += w[a][k] * R[a][k];
        }
      const double tot = warp_transpose_reduce<H>(rs, lane);
#pragma unroll
      for (int k = 0; k < H; ++k) rho[k] = shfl_d(tot, k);
    }
    const bool dual_start = has_w && (opt.dual_init > 0.0);
    if (dual_start) {
      const double zeta0 = has_c ? opt.dual_init : 0.0;
#pragma unroll
      for (int k = 0; k < H; ++k) {
        double mx = 0.0;
#pragma unroll
        for (int a = 0; a < APT; ++a) if (valid[a]) mx = fmax(mx, R[a][k] / rho[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, shfl_xor_d(mx, o));
        nu[k] = mx + opt.dual_init;
        zc[k] = has_c ? zeta0 : 0.0;
      }
#pragma unroll
      for (int a = 0; a < APT; ++a)
#pragma unroll
        for (int k = 0; k < H; ++k) {
          zw[a][k] = valid[a] ? (-R[a][k] / rho[k] + nu[k]) : 0.0;
          zp[a][k] = has_u ? 0.5 * (lam + zeta0) : 0.0;
          zq[a][k] = zp[a][k];
        }
    } else {
#pragma unroll
      for (int k = 0; k < H; ++k) { nu[k] = 1.0; zc[k] = has_c ? opt.mu0 / sc[k] : 0.0; }
#pragma unroll
      for (int a = 0; a < APT; ++a)
#pragma unroll
        for (int k = 0; k < H; ++k) {
          zw[a][k] = (has_w && valid[a]) ? opt.mu0 / w[a][k] : 0.0;
          zp[a][k] = has_u ? opt.mu0 / sp[a][k] : 0.0;
          zq[a][k] = has_u ? opt.mu0 / sq[a][k] : 0.0;
        }
    }
    const double mcount = (has_w ? (double)H * N : 0.0) + (has_u ? 2.0 * H * N : 0.0) + (has_c ? (double)H : 0.0);
    int status = ST_FAILED;
    double gw[APT][H];
    double rp[H];
    for (int it = 1; it <= opt.max_iter + 1; ++it) {
      iters = it;
      // ---- residuals --------------------------------------------------------------------------------
      double rs[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) rs[i] = 0.0;
#pragma unroll
      for (int a = 0; a < APT; ++a)
        if (valid[a]) {
#pragma unroll
          for (int k = 0; k < H; ++k) {
            rs[k] += w[a][k] * R[a][k];
            rs[H + k] += w[a][k];
            double g = 0.0;
            if (has_w) g += w[a][k] * zw[a][k];
            if (has_u) g += sp[a][k] * zp[a][k] + sq[a][k] * zq[a][k];
            rs[2 * H] += g;
          }
        }
      static_assert(2 * H + 1 <= 32, "H too large for the residual batch");
      const double tot = warp_transpose_reduce<2 * H + 1>(rs, lane);
      double pres = 0.0, gap = shfl_d(tot, 2 * H);
#pragma unroll
      for (int k = 0; k < H; ++k) {
        rho[k] = shfl_d(tot, k);
        rp[k] = shfl_d(tot, H + k) - 1.0;
        pres = fmax(pres, fabs(rp[k]));
        if (has_c) gap += sc[k] * zc[k];
      }
      double dres = 0.0;
#pragma unroll
      for (int a = 0; a < APT; ++a) {
#pragma unroll
        for (int k = 0; k < H; ++k) {
          gw[a][k] = -R[a][k] / rho[k];
          if (!valid[a]) continue;
          const double yk = zp[a][k] - zq[a][k];
          const double yn = (k + 1 < H) ? zp[a][k + 1] - zq[a][k + 1] : 0.0;
          const double rdw = gw[a][k] - zw[a][k] + yk - yn + nu[k];
          dres = fmax(dres, fabs(rdw));
          if (has_u) {
            const double rdu = lam - zp[a][k] - zq[a][k] + (has_c ? zc[k] : 0.0);
            dres = fmax(dres, fabs(rdu));
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dres = fmax(dres, shfl_xor_d(dres, o));
      kkt[0] = pres; kkt[1] = dres; kkt[2] = gap;
      if (!isfinite(dres + gap)) break;
      if (pres < opt.tol && dres < opt.tol_dual && gap < opt.tol) { status = ST_OPTIMAL; break; }
      if (it == opt.max_iter + 1) break;
      const double mu = gap / fmax(mcount, 1.0);
      if (!factorize(rho)) break;
      // ---- predictor ------------------------------------------------------------------------------------
      Dir d;
      double cw[APT][H], cp[APT][H], cq[APT][H], cc[H];
#pragma unroll
      for (int a = 0; a < APT; ++a)
#pragma unroll
        for (int k = 0; k < H; ++k) { cw[a][k] = 0.0; cp[a][k] = 0.0; cq[a][k] = 0.0; }
#pragma unroll
      for (int k = 0; k < H; ++k) cc[k] = 0.0;
      double sigma = 0.0;
      if (mcount > 0.0) {
        newton(gw, rp, cw, cp, cq, cc, d);
        double aa, ab;
        max_step(d, rho, allow_short, aa, ab);
        double g2 = 0.0;
#pragma unroll
        for (int a = 0; a < APT; ++a)
          if (valid[a]) {
#pragma unroll
            for (int k = 0; k < H; ++k) {
              if (has_w) g2 += (w[a][k] + aa * d.dw[a][k]) * (zw[a][k] + ab * d.dzw[a][k]);
              if (has_u) g2 += (sp[a][k] + aa * d.dsp[a][k]) * (zp[a][k] + ab * d.dzp[a][k]) +
                               (sq[a][k] + aa * d.dsq[a][k]) * (zq[a][k] + ab * d.dzq[a][k]);
            }
          }
        g2 = warp_sum(g2);
        if (has_c) {
#pragma unroll
          for (int k = 0; k < H; ++k) g2 += (sc[k] + aa * d.dsc[k]) * (zc[k] + ab * d.dzc[k]);
        }
        const double ratio = (gap > 0.0) ? fmin(1.0, fmax(g2 / gap, 0.0)) : 0.0;
        sigma = ratio * ratio * ratio;
        const double sm = sigma * mu;
#pragma unroll
        for (int a = 0; a < APT; ++a)
#pragma unroll
          for (int k = 0; k < H; ++k) {
            cw[a][k] = has_w ? sm - d.dw[a][k] * d.dzw[a][k] : 0.0;
            cp[a][k] = has_u ? sm - d.dsp[a][k] * d.dzp[a][k] : 0.0;
            cq[a][k] = has_u ? sm - d.dsq[a][k] * d.dzq[a][k] : 0.0;
          }
#pragma unroll
        for (int k = 0; k < H; ++k) cc[k] = has_c ? sm - d.dsc[k] * d.dzc[k] : 0.0;
      }
      // ---- corrector ------------------------------------------------------------------------------------
      newton(gw, rp, cw, cp, cq, cc, d);
      double a_ = 1.0, b_ = 1.0;
      if (mcount > 0.0 || allow_short) {
        max_step(d, rho, allow_short, a_, b_);
        a_ = fmin(1.0, opt.step_frac * a_); b_ = fmin(1.0, opt.step_frac * b_);
      }
#pragma unroll
      for (int a = 0; a < APT; ++a)
        if (valid[a]) {
#pragma unroll
          for (int k = 0; k < H; ++k) {
            w[a][k] += a_ * d.dw[a][k];
            if (has_w) zw[a][k] += b_ * d.dzw[a][k];
            if (has_u) {
              sp[a][k] += a_ * d.dsp[a][k]; sq[a][k] += a_ * d.dsq[a][k];
              zp[a][k] += b_ * d.dzp[a][k]; zq[a][k] += b_ * d.dzq[a][k];
            }
          }
        }
#pragma unroll
      for (int k = 0; k < H; ++k) {
        nu[k] += b_ * d.dnu[k];
        if (has_c) { sc[k] += a_ * d.dsc[k]; zc[k] += b_ * d.dzc[k]; }
      }
    }
    if (status != ST_OPTIMAL && isfinite(kkt[1] + kkt[2]) && kkt[0] < 1e-8 && kkt[1] < 1e-6 && kkt[2] < 1e-8)
      status = ST_INACCURATE;
    if (status == ST_FAILED) hold(w0);
    return status;
  }

  __device__ __forceinline__ void hold(const double (&w0)[APT]) {
#pragma unroll
    for (int a = 0; a < APT; ++a)
#pragma unroll
      for (int k = 0; k < H; ++k) w[a][k] = w0[a];
  }

  // maximised objective of mpc.py:104 for the plan in w[][] (fp64)
  __device__ double objective(const double (&w0)[APT]) const {
    double rs[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) rs[i] = 0.0;
#pragma unroll
    for (int a = 0; a < APT; ++a)
      if (valid[a]) {
#pragma unroll
        for (int k = 0; k < H; ++k) {
          rs[k] += w[a][k] * R[a][k];
          rs[H] += fabs(w[a][k] - ((k == 0) ? w0[a] : w[a][k - 1]));
        }
      }
    const double tot = warp_transpose_reduce<H + 1>(rs, lane);
    double val = -lam * shfl_d(tot, H);
#pragma unroll
    for (int k = 0; k < H; ++k) val += log(shfl_d(tot, k));
    return val;
  }
};

}  // namespace kmpc
