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
// Memory placement (the solver is latency/issue bound, not HBM bound): lane l owns assets l, l+32, ...
// (APT per lane).  The iterate (w, slacks, duals, R) lives in registers; the factorisation of the current
// iterate (Green factors, barrier weights) and the search directions live in the warp's private slice of
// shared memory, laid out [array][stage][slot] so that every access is conflict free.  H is a compile-time
// constant: every loop over stages is unrolled and every register array is statically indexed.
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
  // step_frac / dual_init tuned on the 1.0 M-decision config-2 replay and a random instance mix (N 2..64, H 1..5,
  // lam 0..0.1, tau 0..1): (0.995, 3e-3) -> (0.9999, 1e-3) takes 9.35 -> 8.00 iterations per decision at the same
  // failure rate (3 fallbacks per million), 9.8 -> 8.9 on the mix with zero failures
  o.tol = 1e-10; o.tol_dual = 1e-8; o.delta = 1e-5; o.step_frac = 0.9999; o.mu0 = 1e-3; o.dual_init = 1e-3;
  o.max_iter = 100;
  return o;
}

// Acceptance of an iterate the iteration could not push to the tolerances (iteration cap, breakdown of the border
// factorisation once the barrier weights span > 20 decades): "optimal_inaccurate" (mpc.py:113 uses such weights)
// when it is primal feasible, the gap has collapsed and the dual residual is at the level a first-order reference
// solver stops at (SCS eps 1e-4).  Measured on the config-2 replay: every such iterate is within 5e-7 relative of
// the optimal objective and 4e-4 of the optimal first-stage weights, whereas holding the weights (the fallback) is
// 0.1 away.
constexpr double kLoosePres = 1e-8, kLooseDres = 1e-4, kLooseGap = 1e-7;
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
        double lo = (i < NV) ? v[i] : 0.0;
        double hi = (i + o < NV) ? v[i + o] : 0.0;
        double send = up ? lo : hi;
        double keep = up ? hi : lo;
        v[i] = keep + shfl_xor_d(send, o);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] += shfl_xor_d(v[i], o);
    }
  }
  return v[0];
}

template <int H, int APT, int NS>
struct WarpIpm {
  static constexpr int NB = 3 * H;          // border size (R-rows, budget rows, cap rows)
  static constexpr int SLOTS = NS;          // slots per row: N <= NS <= 32*APT (compile-time so that every
                                            // shared-memory access has an immediate offset)
  static_assert(NS <= 32 * APT && NS > 32 * (APT - 1), "slot stride does not match assets-per-lane");
  // shared-memory arrays, each [H][SLOTS]: Green factors, barrier weights, search direction, iterate
  enum : int { QL, TL, QR, TR, GJJ, VD, FL, FR, IE, DW, DSP, DSQ, DZW, DZP, DZQ,
               RR, ZW, WW, SP, SQ, ZP, ZQ, IW, ISP, ISQ, NARR };
  // warp-uniform per-stage scalars, also in shared memory (broadcast reads; every lane writes the same value)
  enum : int { NU, SC, ZC, DNU, DSC, DZC, IRHO, ISC, RHO, RP, CC, YR, YN, YC, NUNI };
  static constexpr int SMEM_DOUBLES = NARR * H * SLOTS + NB * NB + NB + NUNI * H;
  static_assert(NB <= 32, "border must fit one entry per lane");
  static_assert(2 * H + 1 <= 32, "H too large for the residual batch");

  double* sm;                               // this warp's shared-memory slice
  double* Ksm;                              // [NB*NB] lower Cholesky factor (row-major), then [NB] 1/diag
  int lane, nassets;
  bool has_w, has_u, has_c;
  double lam, tau, delta;
  int nb;                                   // active border size: 2H or 3H

  __device__ __forceinline__ void bind(double* smem_slice, int lane_, int n_assets) {
    sm = smem_slice; Ksm = smem_slice + NARR * H * SLOTS; lane = lane_; nassets = n_assets;
  }
  __device__ __forceinline__ double& U(int arr, int k) const { return Ksm[NB * NB + NB + arr * H + k]; }
  __device__ __forceinline__ bool ok(int a) const { return lane + 32 * a < nassets; }
  // Only lanes that own asset (a) may touch F(.,.,a): slots >= NS do not exist.  Every use is guarded by ok(a).
  __device__ __forceinline__ double& F(int arr, int k, int a) const { return sm[(arr * H + k) * SLOTS + a * 32 + lane]; }
  __device__ __forceinline__ double dP(int k, int a) const { return F(ZP, k, a) * F(ISP, k, a); }
  __device__ __forceinline__ double dQ(int k, int a) const { return F(ZQ, k, a) * F(ISQ, k, a); }
  __device__ __forceinline__ double dW0(int k, int a) const { return (has_w && ok(a)) ? F(ZW, k, a) * F(IW, k, a) : 0.0; }
  __device__ __forceinline__ double phi(int k, int a) const { return has_u ? (dQ(k, a) - dP(k, a)) * F(IE, k, a) : 0.0; }

  // Green's function column j (runtime) of asset a: G[l] = potential of node l, D[l] = drop across edge l, for
  // a unit current injected at node j.  Register arrays are indexed statically; j only enters predicates.
  __device__ __forceinline__ void green_col(int a, int j, double (&G)[H], double (&D)[H]) const {
    const double gjj = F(GJJ, j, a);
    double v = gjj;
#pragma unroll
    for (int l = 0; l < H; ++l) {
      G[l] = 0.0; D[l] = 0.0;
      if (l == j) G[l] = gjj;
      if (l > j) { D[l] = -v * F(QR, l, a); v *= F(TR, l, a); G[l] = v; }
    }
    v = gjj;
#pragma unroll
    for (int l = H - 1; l >= 0; --l) {
      if (l <= j) {
        D[l] = v * F(QL, l, a); v *= F(TL, l, a);
        if (l >= 1) G[(l >= 1) ? l - 1 : 0] = v;
      }
    }
  }
  // DD[l]: drop across edge l for a unit dipole across edge k (+1 at node k, -1 at node k-1), k runtime
  __device__ __forceinline__ void dipole_col(int a, int k, double (&DD)[H]) const {
    const double V = F(VD, k, a);
    double v = V * F(FL, k, a);
#pragma unroll
    for (int l = 0; l < H; ++l) {
      DD[l] = 0.0;
      if (l == k) DD[l] = V;
      if (l > k) { DD[l] = -v * F(QR, l, a); v *= F(TR, l, a); }
    }
    v = -V * F(FR, k, a);
#pragma unroll
    for (int l = H - 1; l >= 0; --l) {
      if (l < k) { DD[l] = v * F(QL, l, a); v *= F(TL, l, a); }
    }
  }

  // M0^{-1} (g_w, g_u): dw, dd through the Green's functions; pg = phi * g_u (dipole strengths)
  __device__ __forceinline__ void m0_apply(int a, const double (&gwv)[H], const double (&pg)[H],
                                           double (&dw)[H], double (&dd)[H]) const {
#pragma unroll
    for (int k = 0; k < H; ++k) { dw[k] = 0.0; dd[k] = 0.0; }
#pragma unroll
    for (int j = 0; j < H; ++j) {
      double G[H], D[H];
      green_col(a, j, G, D);
      const double g = gwv[j];
      double acc = 0.0;                     // sum_k D[k,j] * pg[k]  (potential of node j from the dipoles)
#pragma unroll
      for (int l = 0; l < H; ++l) {
        dw[l] += G[l] * g;
        dd[l] += D[l] * g;
        acc += D[l] * pg[l];
      }
      dw[j] -= acc;
    }
    if (has_u) {
#pragma unroll
      for (int k = 0; k < H; ++k) {
        double DD[H];
        dipole_col(a, k, DD);
#pragma unroll
        for (int l = 0; l < H; ++l) dd[l] -= DD[l] * pg[k];
      }
    }
  }

  // Build the Green factors of every owned asset (to shared memory) and the border matrix K; Cholesky.
  // Returns false on a non-positive pivot.
  __device__ __forceinline__ bool factorize() {
#pragma unroll 1
    for (int a = 0; a < APT; ++a) {
      if (!ok(a)) continue;
      const bool va = true;
      double e[H], hLv[H], hRv[H], ad[H];
#pragma unroll
      for (int k = 0; k < H; ++k) {
        const double iw = 1.0 / F(WW, k, a);
        F(IW, k, a) = iw;
        const double dw0 = (has_w && va) ? F(ZW, k, a) * iw : 0.0;
        ad[k] = dw0 + delta;
        if (has_u) {
          const double isp = 1.0 / F(SP, k, a), isq = 1.0 / F(SQ, k, a);
          F(ISP, k, a) = isp; F(ISQ, k, a) = isq;
          const double dp = F(ZP, k, a) * isp, dq = F(ZQ, k, a) * isq;
          const double E = dp + dq + delta;
          const double ie = 1.0 / E;
          F(IE, k, a) = ie;
          e[k] = (4.0 * dp * dq + 2.0 * delta * (dp + dq) + delta * delta) * ie;
        } else {
          F(IE, k, a) = 1.0; e[k] = 0.0;
        }
      }
      // left sweep: hL[k] = conductance to ground seen at node k leftwards incl. ad[k]
#pragma unroll
      for (int l = 0; l < H; ++l) {
        double ql = 1.0, tl = 0.0;
        if (l > 0) {
          const double inv = 1.0 / (e[l] + hLv[(l > 0) ? l - 1 : 0]);
          ql = hLv[(l > 0) ? l - 1 : 0] * inv; tl = e[l] * inv;
        }
        F(QL, l, a) = ql; F(TL, l, a) = tl;
        hLv[l] = ad[l] + e[l] * ql;
      }
      // right sweep: hR[k] incl. ad[k]; qR/tR indexed by the edge entering node l from the left
      hRv[H - 1] = ad[H - 1];
      double qr[H];
      qr[0] = 0.0;
      F(QR, 0, a) = 0.0; F(TR, 0, a) = 0.0;
#pragma unroll
      for (int l = H - 1; l >= 1; --l) {
        const double inv = 1.0 / (e[l] + hRv[l]);
        qr[l] = hRv[l] * inv;
        F(QR, l, a) = qr[l]; F(TR, l, a) = e[l] * inv;
        hRv[l - 1] = ad[l - 1] + e[l] * qr[l];
      }
#pragma unroll
      for (int j = 0; j < H; ++j) {
        const int jn = (j + 1 < H) ? j + 1 : 0;
        const double gR = (j + 1 < H) ? e[jn] * qr[jn] : 0.0;
        F(GJJ, j, a) = 1.0 / (hLv[j] + gR);
        double fl = 1.0, fr = 0.0;
        if (j > 0) {
          const double inv = 1.0 / (hLv[(j > 0) ? j - 1 : 0] + hRv[j]);
          fl = hLv[(j > 0) ? j - 1 : 0] * inv; fr = hRv[j] * inv;
        }
        F(FL, j, a) = fl; F(FR, j, a) = fr;
        F(VD, j, a) = 1.0 / (e[j] + hRv[j] * fl);
      }
    }
    __syncwarp();
    // K assembly, one batch per column stage j.  Border rows: [0,H) = Rt, [H,2H) = 1t, [2H,3H) = et.
    // Batch j holds, for all row stages l:  (1t_l,Rt_j) [H]  and for l >= j: (Rt_l,Rt_j) [H]  (1t_l,1t_j) [H]
    //                then the cap entries:  (et_l,Rt_j) [H]  (et_l,1t_j) [H]  and for l >= j: (et_l,et_j) [H]
    // For H > 5 the batch is split in two halves of <= 32 entries.
    constexpr bool kSplit = (6 * H > 32);
    constexpr int kHalves = kSplit ? 2 : 1;
#pragma unroll 1
    for (int jh = 0; jh < H * kHalves; ++jh) {
      const int j = jh / kHalves, half = jh - j * kHalves;
      double c[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) c[i] = 0.0;
      const bool doG = !kSplit || half == 0;       // entries built from G: 1R, RR, 11
      const bool doE = (!kSplit || half == 1) && has_c;   // entries involving cap rows: eR, e1, ee
      constexpr int offE = kSplit ? 0 : 3 * H;     // position of the cap entries inside c[]
#pragma unroll 1
      for (int a = 0; a < APT; ++a) {
        if (!ok(a)) continue;
        double G[H], D[H];
        green_col(a, j, G, D);
        const double Rj = F(RR, j, a);
        if (doG) {
#pragma unroll
          for (int l = 0; l < H; ++l) {
            c[l] += Rj * G[l];                                   // (1t_l, Rt_j)
            if (l >= j) {
              c[H + l] += F(RR, l, a) * Rj * G[l];               // (Rt_l, Rt_j)
              c[2 * H + l] += G[l];                              // (1t_l, 1t_j)
            }
          }
        }
        if (doE) {
          double DD[H];
          dipole_col(a, j, DD);
          const double phj = phi(j, a);
#pragma unroll
          for (int l = 0; l < H; ++l) {
            const double phl = phi(l, a);
            const double dm = -phl * D[l];
            c[offE + l] += dm * Rj;                              // (et_l, Rt_j)
            c[offE + H + l] += dm;                               // (et_l, 1t_j)
            if (l >= j) {
              double v = phl * phj * DD[l];
              if (l == j) v += F(IE, l, a);
              c[offE + 2 * H + l] += v;                          // (et_l, et_j)
            }
          }
        }
      }
      const double tot = warp_transpose_reduce<(kSplit ? 3 * H : 6 * H)>(c, lane);
      // scatter: lane e holds entry e of this batch
      int t = lane;
      bool capPart = kSplit ? (half == 1) : false;
      if (!kSplit && t >= 3 * H) { t -= 3 * H; capPart = true; }
      if (lane < (kSplit ? 3 * H : 6 * H)) {
        const int grp = t / H, l = t - grp * H;
        int r = -1, cc = -1;
        if (!capPart) {
          if (grp == 0) { r = H + l; cc = j; }
          else if (grp == 1 && l >= j) { r = l; cc = j; }
          else if (grp == 2 && l >= j) { r = H + l; cc = H + j; }
        } else if (has_c) {
          if (grp == 0) { r = 2 * H + l; cc = j; }
          else if (grp == 1) { r = 2 * H + l; cc = H + j; }
          else if (grp == 2 && l >= j) { r = 2 * H + l; cc = 2 * H + j; }
        }
        if (r >= 0) Ksm[r * NB + cc] = tot;
      }
    }
    __syncwarp();
    if (lane == 0) {          // static indices only: a lane-indexed pick would push rho/sc/zc to local memory
#pragma unroll
      for (int k = 0; k < H; ++k) {
        Ksm[k * NB + k] += U(RHO, k) * U(RHO, k);                                   // 1/beta_k
        if (has_c) Ksm[(2 * H + k) * NB + 2 * H + k] += U(SC, k) / U(ZC, k);
      }
    }
    __syncwarp();
    // warp Cholesky (lower, in place); lane i owns row i
    bool pd = true;
#pragma unroll 1
    for (int j = 0; j < nb; ++j) {
      const double djj = Ksm[j * NB + j];
      if (!(djj > 0.0)) { pd = false; break; }
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
    return pd;
  }

  // y <- K^{-1} t ; lane i holds t_i on entry and y_i on exit (i < nb)
  __device__ __forceinline__ double k_solve(double t) const {
#pragma unroll 1
    for (int j = 0; j < nb; ++j) {                 // forward: L y = t
      const double yj = shfl_d(t, j) * Ksm[NB * NB + j];
      if (lane == j) t = yj;
      if (lane > j && lane < nb) t -= Ksm[lane * NB + j] * yj;
    }
#pragma unroll 1
    for (int j = nb - 1; j >= 0; --j) {            // backward: L' x = y
      const double xj = shfl_d(t, j) * Ksm[NB * NB + j];
      if (lane == j) t = xj;
      if (lane < j) t -= Ksm[j * NB + lane] * xj;
    }
    return t;
  }

  // right-hand side of one asset:  rhs_x = -grad f - A' nu - G'(c/s);  c-terms come from shared memory
  // (they alias the DZ* arrays, see solve()) when use_c, else they are zero (predictor).
  __device__ __forceinline__ void build_rhs(int a, bool use_c, double (&g_w)[H], double (&g_u)[H]) const {
    double tq[H];
#pragma unroll
    for (int k = 0; k < H; ++k) {
      double gwk = F(RR, k, a) * U(IRHO, k) - U(NU, k);              // -grad_w f - nu
      if (has_w && use_c) gwk += F(DZW, k, a) * F(IW, k, a);
      double guk = 0.0;
      tq[k] = 0.0;
      if (has_u) {
        double a1 = 0.0, a2 = 0.0;
        if (use_c) { a1 = F(DZP, k, a) * F(ISP, k, a); a2 = F(DZQ, k, a) * F(ISQ, k, a); }
        tq[k] = a1 - a2;
        guk = -lam + a1 + a2;
        if (has_c) guk -= U(CC, k) * U(ISC, k);
      }
      g_w[k] = gwk; g_u[k] = guk;
    }
#pragma unroll
    for (int k = 0; k < H; ++k) {
      g_w[k] -= tq[k];
      if (k + 1 < H) g_w[k] += tq[(k + 1 < H) ? k + 1 : 0];
    }
  }

  // One Newton solve with complementarity targets c (sigma*mu - corrector products; zero when !use_c).
  // Writes the direction to shared memory (DW, DSP, DSQ, DZW, DZP, DZQ) and dnu/dsc/dzc.
  // Two passes over the assets share one code instance: pass 0 accumulates t = V' M0^{-1} g, pass 1 applies
  // dx = M0^{-1}(g - V y).
  __device__ __forceinline__ void newton(bool use_c) {
    double tv[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) tv[i] = 0.0;
#pragma unroll
    for (int k = 0; k < H; ++k) { U(YR, k) = 0.0; U(YN, k) = 0.0; U(YC, k) = 0.0; }
    __syncwarp();
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      if (pass == 1) {
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < H; ++k) tv[H + k] += U(RP, k);            // t[H+k] = sum dw0 - q, q = -rp
        }
        const double t = warp_transpose_reduce<NB>(tv, lane);
        const double y = k_solve(t);
#pragma unroll
        for (int k = 0; k < H; ++k) {
          U(YR, k) = shfl_d(y, k);
          U(YN, k) = shfl_d(y, H + k);
          U(YC, k) = has_c ? shfl_d(y, 2 * H + k) : 0.0;
          U(DNU, k) = U(YN, k);
          if (has_c) {
            U(DSC, k) = -U(YC, k) * U(SC, k) / U(ZC, k);
            U(DZC, k) = (U(CC, k) * U(ISC, k) - U(ZC, k)) + U(YC, k);
          } else { U(DSC, k) = 0.0; U(DZC, k) = 0.0; }
        }
      }
#pragma unroll 1
      for (int a = 0; a < APT; ++a) {
        if (!ok(a)) continue;
        double g_w[H], g_u[H], pg[H], dw[H], dd[H];
        build_rhs(a, use_c, g_w, g_u);
#pragma unroll
        for (int k = 0; k < H; ++k) {
          g_w[k] -= U(YR, k) * F(RR, k, a) + U(YN, k);                // zeros in pass 0
          g_u[k] -= U(YC, k);                                      // geff
          pg[k] = phi(k, a) * g_u[k];
        }
        m0_apply(a, g_w, pg, dw, dd);
        if (pass == 0) {
#pragma unroll
          for (int k = 0; k < H; ++k) {
            tv[k] += F(RR, k, a) * dw[k];
            tv[H + k] += dw[k];
            if (has_c) tv[2 * H + k] += g_u[k] * F(IE, k, a) - phi(k, a) * dd[k];   // du0
          }
        } else {
#pragma unroll
          for (int k = 0; k < H; ++k) {
            const double cwv = (use_c && has_w) ? F(DZW, k, a) : 0.0;
            const double cpv = (use_c && has_u) ? F(DZP, k, a) : 0.0;
            const double cqv = (use_c && has_u) ? F(DZQ, k, a) : 0.0;
            F(DW, k, a) = dw[k];
            F(DZW, k, a) = has_w ? (cwv * F(IW, k, a) - F(ZW, k, a)) - dW0(k, a) * dw[k] : 0.0;
            if (has_u) {
              const double dp = dP(k, a), dq = dQ(k, a), ie = F(IE, k, a);
              const double dsp_ = (g_u[k] - (2.0 * dq + delta) * dd[k]) * ie;
              const double dsq_ = (g_u[k] + (2.0 * dp + delta) * dd[k]) * ie;
              F(DSP, k, a) = dsp_; F(DSQ, k, a) = dsq_;
              F(DZP, k, a) = (cpv * F(ISP, k, a) - F(ZP, k, a)) - dp * dsp_;
              F(DZQ, k, a) = (cqv * F(ISQ, k, a) - F(ZQ, k, a)) - dq * dsq_;
            } else {
              F(DSP, k, a) = 0.0; F(DSQ, k, a) = 0.0; F(DZP, k, a) = 0.0; F(DZQ, k, a) = 0.0;
            }
          }
        }
      }
    }
  }

  // largest steps keeping slacks (ap) and duals (ad) positive
  __device__ __forceinline__ void max_step(bool allow_short, double& ap, double& ad) const {
    double p = 1.0, q = 1.0;
    // a0 <- min(a0, -v/dv) for dv < 0; the division is executed only when it lowers the bound
    auto lim = [](double v, double dv, double a0) {
      if (dv < 0.0 && fma(a0, dv, v) < 0.0) a0 = fmin(a0, -v / dv);
      return a0;
    };
    double dr[32];
    if (allow_short) {
#pragma unroll
      for (int i = 0; i < 32; ++i) dr[i] = 0.0;
    }
#pragma unroll 1
    for (int a = 0; a < APT; ++a) {
      if (!ok(a)) continue;
#pragma unroll
      for (int k = 0; k < H; ++k) {
        if (has_w) { p = lim(F(WW, k, a), F(DW, k, a), p); q = lim(F(ZW, k, a), F(DZW, k, a), q); }
        if (has_u) {
          p = lim(F(SP, k, a), F(DSP, k, a), p); p = lim(F(SQ, k, a), F(DSQ, k, a), p);
          q = lim(F(ZP, k, a), F(DZP, k, a), q); q = lim(F(ZQ, k, a), F(DZQ, k, a), q);
        }
        if (allow_short) dr[k] += F(DW, k, a) * F(RR, k, a);
      }
    }
    if (allow_short) {       // keep the log argument positive
      const double tot = warp_transpose_reduce<H>(dr, lane);
#pragma unroll
      for (int k = 0; k < H; ++k) p = lim(U(RHO, k), shfl_d(tot, k), p);
    }
    if (has_c) {
#pragma unroll
      for (int k = 0; k < H; ++k) { p = lim(U(SC, k), U(DSC, k), p); q = lim(U(ZC, k), U(DZC, k), q); }
    }
    // exact fp64 min over the warp (keeps the iterates identical to the oracle's)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { p = fmin(p, shfl_xor_d(p, o)); q = fmin(q, shfl_xor_d(q, o)); }
    ap = p; ad = q;
  }

  // Solve.  On entry the RR array (gross returns) is filled and bind() was called.  w0[a] = current weights of
  // the owned assets.  Returns status; the WW array holds the plan (or tile(w0) on failure), kkt = (pres, dres, gap).
  __device__ __forceinline__ int solve(const double (&w0)[APT], int N, double lam_, double tau_, bool allow_short,
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
      if (ok(a)) {
        if (!isfinite(w0[a])) bad = 1;
#pragma unroll
        for (int k = 0; k < H; ++k) if (!(isfinite(F(RR, k, a)) && F(RR, k, a) > 0.0)) bad = 1;
      }
    if (__any_sync(kFull, bad)) { hold(w0); return ST_NONFINITE; }
    // ---- initial point (oracle/mpc_oracle.py::_initial_point) ----------------------------------------
    double base[APT], sb = 0.0;
#pragma unroll
    for (int a = 0; a < APT; ++a) {
      base[a] = ok(a) ? (allow_short ? w0[a] : fmax(w0[a], 0.0)) : 0.0;
      sb += base[a];
    }
    sb = warp_sum(sb);
    const double invN = 1.0 / (double)N;
    const double eps = (tau <= 0.0) ? 0.1 : fmin(0.1, tau / 8.0);
    double absd0 = 0.0;
#pragma unroll
    for (int a = 0; a < APT; ++a) {
      const double b = (sb > 0.0) ? base[a] / sb : invN;
      const double w1 = (1.0 - eps) * b + eps * invN;
      if (ok(a)) {
#pragma unroll
        for (int k = 0; k < H; ++k) F(WW, k, a) = w1;
        absd0 += fabs(w1 - w0[a]);
      }
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
        if (!ok(a)) continue;
        const double d0 = F(WW, 0, a) - w0[a];
        const double u0 = fabs(d0) + dl0;
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const double dk = (k == 0) ? d0 : 0.0;
          const double uk = (k == 0) ? u0 : dlk;
          F(SP, k, a) = uk - dk; F(SQ, k, a) = uk + dk;
        }
        su0 += u0;
      }
      su0 = warp_sum(su0);
#pragma unroll
      for (int k = 0; k < H; ++k) U(SC, k) = has_c ? (tau - ((k == 0) ? su0 : dlk * N)) : 1.0;
    } else {
#pragma unroll
      for (int a = 0; a < APT; ++a)
        if (ok(a)) {
#pragma unroll
          for (int k = 0; k < H; ++k) { F(SP, k, a) = 1.0; F(SQ, k, a) = 1.0; }
        }
#pragma unroll
      for (int k = 0; k < H; ++k) U(SC, k) = 1.0;
    }
    {
      double rs[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) rs[i] = 0.0;
#pragma unroll
      for (int a = 0; a < APT; ++a)
        if (ok(a)) {
#pragma unroll
          for (int k = 0; k < H; ++k) rs[k] += F(WW, k, a) * F(RR, k, a);
        }
      const double tot = warp_transpose_reduce<H>(rs, lane);
#pragma unroll
      for (int k = 0; k < H; ++k) U(RHO, k) = shfl_d(tot, k);
    }
    const bool dual_start = has_w && (opt.dual_init > 0.0);
    if (dual_start) {
      const double zeta0 = has_c ? opt.dual_init : 0.0;
#pragma unroll
      for (int k = 0; k < H; ++k) {
        double mx = 0.0;
#pragma unroll
        for (int a = 0; a < APT; ++a) if (ok(a)) mx = fmax(mx, F(RR, k, a) / U(RHO, k));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, shfl_xor_d(mx, o));
        U(NU, k) = mx + opt.dual_init;
        U(ZC, k) = has_c ? zeta0 : 0.0;
      }
#pragma unroll
      for (int a = 0; a < APT; ++a)
        if (ok(a)) {
#pragma unroll
          for (int k = 0; k < H; ++k) {
            F(ZW, k, a) = -F(RR, k, a) / U(RHO, k) + U(NU, k);
            F(ZP, k, a) = has_u ? 0.5 * fmax(lam + zeta0, opt.dual_init) : 0.0;   // floor: see oracle
            F(ZQ, k, a) = F(ZP, k, a);
          }
        }
    } else {
#pragma unroll
      for (int k = 0; k < H; ++k) { U(NU, k) = 1.0; U(ZC, k) = has_c ? opt.mu0 / U(SC, k) : 0.0; }
#pragma unroll
      for (int a = 0; a < APT; ++a)
        if (ok(a)) {
#pragma unroll
          for (int k = 0; k < H; ++k) {
            F(ZW, k, a) = has_w ? opt.mu0 / F(WW, k, a) : 0.0;
            F(ZP, k, a) = has_u ? opt.mu0 / F(SP, k, a) : 0.0;
            F(ZQ, k, a) = has_u ? opt.mu0 / F(SQ, k, a) : 0.0;
          }
        }
    }
    const double mcount = (has_w ? (double)H * N : 0.0) + (has_u ? 2.0 * H * N : 0.0) + (has_c ? (double)H : 0.0);
    int status = ST_FAILED;
#pragma unroll 1
    for (int it = 1; it <= opt.max_iter + 1; ++it) {
      iters = it;
      // ---- residuals --------------------------------------------------------------------------------
      double rs[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) rs[i] = 0.0;
#pragma unroll 1
      for (int a = 0; a < APT; ++a) {
        if (!ok(a)) continue;
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const double wk = F(WW, k, a);
          rs[k] += wk * F(RR, k, a);
          rs[H + k] += wk;
          double g = 0.0;
          if (has_w) g += wk * F(ZW, k, a);
          if (has_u) g += F(SP, k, a) * F(ZP, k, a) + F(SQ, k, a) * F(ZQ, k, a);
          rs[2 * H] += g;
        }
      }
      const double tot = warp_transpose_reduce<2 * H + 1>(rs, lane);
      double pres = 0.0, gap = shfl_d(tot, 2 * H);
#pragma unroll
      for (int k = 0; k < H; ++k) {
        U(RHO, k) = shfl_d(tot, k);
        U(IRHO, k) = 1.0 / U(RHO, k);
        U(ISC, k) = has_c ? 1.0 / U(SC, k) : 0.0;
        U(RP, k) = shfl_d(tot, H + k) - 1.0;
        pres = fmax(pres, fabs(U(RP, k)));
        if (has_c) gap += U(SC, k) * U(ZC, k);
      }
      double dres = 0.0;
#pragma unroll 1
      for (int a = 0; a < APT; ++a) {
        if (!ok(a)) continue;
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const int kn = (k + 1 < H) ? k + 1 : 0;
          const double yk = F(ZP, k, a) - F(ZQ, k, a);
          const double yn = (k + 1 < H) ? F(ZP, kn, a) - F(ZQ, kn, a) : 0.0;
          const double rdw = -F(RR, k, a) * U(IRHO, k) - F(ZW, k, a) + yk - yn + U(NU, k);
          dres = fmax(dres, fabs(rdw));
          if (has_u) {
            const double rdu = lam - F(ZP, k, a) - F(ZQ, k, a) + (has_c ? U(ZC, k) : 0.0);
            dres = fmax(dres, fabs(rdu));
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dres = fmax(dres, shfl_xor_d(dres, o));
      kkt[0] = pres; kkt[1] = dres; kkt[2] = gap;
      if (!isfinite(dres + gap)) break;
      if (pres < opt.tol && dres < opt.tol_dual && gap < opt.tol) { status = ST_OPTIMAL; break; }
      // flat directions (curvature << delta): the dual residual crawls at ~delta*|dx| while the gap has long
      // collapsed; the objective is converged -> "optimal_inaccurate" instead of iterating into round-off
      if (pres < opt.tol && gap < 1e-6 * opt.tol && dres < 1e-6) { status = ST_INACCURATE; break; }
      if (it == opt.max_iter + 1) break;
      const double mu = gap / fmax(mcount, 1.0);
      // endgame: primal residual and gap converged, only the dual residual along flat directions is left ->
      // shrink the proximal term so the Newton step is no longer damped there
      if (pres < opt.tol && gap < opt.tol) delta = fmax(0.3 * delta, 1e-9);
      __syncwarp();
      if (!factorize()) break;
      // ---- predictor (phase 0) and corrector (phase 1) share one code instance ---------------------------
#pragma unroll
      for (int k = 0; k < H; ++k) U(CC, k) = 0.0;
#pragma unroll 1
      for (int phase = (mcount > 0.0 ? 0 : 1); phase < 2; ++phase) {
        const bool use_c = (phase == 1) && (mcount > 0.0);
        __syncwarp();
        newton(use_c);
        double aa = 1.0, ab = 1.0;
        if (mcount > 0.0 || allow_short) max_step(allow_short, aa, ab);
        if (phase == 0) {
          double g2 = 0.0;
#pragma unroll 1
          for (int a = 0; a < APT; ++a) {
            if (!ok(a)) continue;
#pragma unroll
            for (int k = 0; k < H; ++k) {
              if (has_w) g2 += (F(WW, k, a) + aa * F(DW, k, a)) * (F(ZW, k, a) + ab * F(DZW, k, a));
              if (has_u) g2 += (F(SP, k, a) + aa * F(DSP, k, a)) * (F(ZP, k, a) + ab * F(DZP, k, a)) +
                               (F(SQ, k, a) + aa * F(DSQ, k, a)) * (F(ZQ, k, a) + ab * F(DZQ, k, a));
            }
          }
          g2 = warp_sum(g2);
          if (has_c) {
#pragma unroll
            for (int k = 0; k < H; ++k) g2 += (U(SC, k) + aa * U(DSC, k)) * (U(ZC, k) + ab * U(DZC, k));
          }
          const double ratio = (gap > 0.0) ? fmin(1.0, fmax(g2 / gap, 0.0)) : 0.0;
          const double sigma = ratio * ratio * ratio;
          const double smu = sigma * mu;
          const double dmp = fmin(1.0, fmin(aa, ab) * (1.0 / kCorrFull));
          // complementarity targets of the corrector, stored in place of the affine dual steps
#pragma unroll 1
          for (int a = 0; a < APT; ++a) {
            if (!ok(a)) continue;
#pragma unroll
            for (int k = 0; k < H; ++k) {
              F(DZW, k, a) = has_w ? smu - dmp * F(DW, k, a) * F(DZW, k, a) : 0.0;
              F(DZP, k, a) = has_u ? smu - dmp * F(DSP, k, a) * F(DZP, k, a) : 0.0;
              F(DZQ, k, a) = has_u ? smu - dmp * F(DSQ, k, a) * F(DZQ, k, a) : 0.0;
            }
          }
#pragma unroll
          for (int k = 0; k < H; ++k) U(CC, k) = has_c ? smu - dmp * U(DSC, k) * U(DZC, k) : 0.0;
        } else {
          const double a_ = fmin(1.0, opt.step_frac * aa), b_ = fmin(1.0, opt.step_frac * ab);
          const double pa = (mcount > 0.0 || allow_short) ? a_ : 1.0, pb = (mcount > 0.0 || allow_short) ? b_ : 1.0;
#pragma unroll 1
          for (int a = 0; a < APT; ++a) {
            if (!ok(a)) continue;
#pragma unroll
            for (int k = 0; k < H; ++k) {
              F(WW, k, a) += pa * F(DW, k, a);
              if (has_w) F(ZW, k, a) += pb * F(DZW, k, a);
              if (has_u) {
                F(SP, k, a) += pa * F(DSP, k, a); F(SQ, k, a) += pa * F(DSQ, k, a);
                F(ZP, k, a) += pb * F(DZP, k, a); F(ZQ, k, a) += pb * F(DZQ, k, a);
              }
            }
          }
#pragma unroll
          for (int k = 0; k < H; ++k) {
            U(NU, k) += pb * U(DNU, k);
            if (has_c) { U(SC, k) += pa * U(DSC, k); U(ZC, k) += pb * U(DZC, k); }
          }
        }
      }
    }
    if (status != ST_OPTIMAL && isfinite(kkt[1] + kkt[2]) && kkt[0] < kLoosePres && kkt[1] < kLooseDres && kkt[2] < kLooseGap)
      status = ST_INACCURATE;
    if (status == ST_FAILED) hold(w0);
    return status;
  }

  __device__ __forceinline__ void hold(const double (&w0)[APT]) {
#pragma unroll
    for (int a = 0; a < APT; ++a)
      if (ok(a)) {
#pragma unroll
        for (int k = 0; k < H; ++k) F(WW, k, a) = w0[a];
      }
  }

  // maximised objective of mpc.py:104 for the plan in the WW array (fp64)
  __device__ __forceinline__ double objective(const double (&w0)[APT]) const {
    double rs[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) rs[i] = 0.0;
#pragma unroll
    for (int a = 0; a < APT; ++a)
      if (ok(a)) {
#pragma unroll
        for (int k = 0; k < H; ++k) {
          rs[k] += F(WW, k, a) * F(RR, k, a);
          rs[H] += fabs(F(WW, k, a) - ((k == 0) ? w0[a] : F(WW, (k == 0) ? 0 : k - 1, a)));
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
