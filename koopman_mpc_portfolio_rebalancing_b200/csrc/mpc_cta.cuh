// CTA-per-problem variant of the fp64 interior-point MPC solver (same algorithm, iteration for iteration, as
// mpc_ipm.cuh and oracle/mpc_oracle.py::solve_structured; see mpc_ipm.cuh for the mathematics).
//
// Why a second layout: the warp-per-problem kernel keeps every per-(asset, stage) quantity of a problem in ONE
// warp's shared-memory slice (~52 KB at N = 50, H = 5), so only 4 warps fit an SM and the kernel is bound by
// dependency latency (profiles/r1_backtest_kernel_v3_rolled.txt).  Here one thread owns ONE (stage k, asset i)
// pair: its slice of the iterate, of the Green's functions and of the search direction lives in ITS registers,
// a block of H * G warps (G = ceil(N / 32)) cooperates on one problem, and shared memory only carries what must
// cross threads (factor sweeps, right-hand sides, the <= 3H border system, reductions).
//
//   thread (k, i):  warp = k * G + i / 32,  lane = i % 32      (all lanes of a warp belong to the same stage)
//
// Per-stage sums (rho_k, budget residual, ...) are warp reductions followed by a G-way combine through shared
// memory; stage-uniform scalars are recomputed identically by every thread of the stage and live in registers.
#pragma once
#include "mpc_ipm.cuh"

namespace kmpc {

template <int H, int G>
struct CtaIpm {
  static constexpr int NW = H * G;             // warps per block
  static constexpr int NT = 32 * NW;
  static constexpr int NSL = 32 * G;           // asset slots per stage
  static constexpr int NB = 3 * H;
  static constexpr int RED = 4;                // values per reduction round
  // shared arrays [H][NSL]
  enum : int { AD, EE, QL, TL, QR, TR, HL, HR, GJJ, VD, FL, FR, XR, XPHI, XGW, XPG, XTQ, NARR };
  static constexpr int OFF_RED = NARR * H * NSL;                 // [2][NW][RED]
  static constexpr int OFF_K = OFF_RED + 2 * NW * RED;            // [NB*NB] + [NB] inverse diagonal
  static constexpr int OFF_T = OFF_K + NB * NB + NB;              // [NB] right-hand side / solution of the border
  static constexpr int OFF_KP = OFF_T + NB;                       // [NW][32] per-warp partial border entries
  static constexpr int OFF_W0 = OFF_KP + NW * 32;                 // [NSL] current weights (all stages read them)
  static constexpr int SMEM_DOUBLES = OFF_W0 + NSL;
  static_assert(NB <= 32 && 6 * H <= 32, "H too large for this layout");

  // ---- registers of thread (k, i) -------------------------------------------------------------------------
  double R, w, sp, sq, zw, zp, zq;             // iterate
  double iw, isp, isq, ie, ph;                 // reciprocals and phi of the current factorisation
  double Gr[H], Dr[H], Dc[H], DDr[H];          // rows/columns of the Green's functions that involve my node/edge
  double dw, dsp, dsq, dzw, dzp, dzq;          // direction (dz* double as complementarity targets, see solve())
  // stage-uniform scalars (identical in every thread of stage k)
  double nu, sc, zc, rho, irho, isc, rp, cc, dnu, dsc, dzc;
  double* sm;
  int k, i, lane, warp, nassets, red_sel;
  bool valid, has_w, has_u, has_c;
  double lam, tau, delta;
  int nb;

  __device__ __forceinline__ void bind(double* smem, int n_assets) {
    sm = smem; nassets = n_assets;
    warp = threadIdx.x >> 5; lane = threadIdx.x & 31;
    k = warp / G; i = (warp - k * G) * 32 + lane;
    valid = i < n_assets; red_sel = 0;
  }
  __device__ __forceinline__ double& X(int arr, int kk, int ii) const { return sm[(arr * H + kk) * NSL + ii]; }
  __device__ __forceinline__ double& Kx(int r, int c) const { return sm[OFF_K + r * NB + c]; }

  // ---- reductions -----------------------------------------------------------------------------------------------
  // NV values per thread -> S = sum over my stage, T = sum over the block.  One __syncthreads per call.
  template <int NV>
  __device__ __forceinline__ void reduce_sum(double (&v)[NV], double (&S)[NV], double (&T)[NV]) {
    static_assert(NV <= RED, "");
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = warp_sum(v[j]);
    double* red = sm + OFF_RED + red_sel * NW * RED;
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < NV; ++j) red[warp * RED + j] = v[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NV; ++j) { S[j] = 0.0; T[j] = 0.0; }
#pragma unroll
    for (int wv = 0; wv < NW; ++wv) {
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const double x = red[wv * RED + j];
        T[j] += x;
        if (wv / G == k) S[j] += x;
      }
    }
    red_sel ^= 1;
  }
  // two values per thread -> block-wide (fmin, fmin) or (fmax, fmax)
  template <bool IS_MIN>
  __device__ __forceinline__ void reduce_ext2(double& a, double& b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double a2 = shfl_xor_d(a, o), b2 = shfl_xor_d(b, o);
      a = IS_MIN ? fmin(a, a2) : fmax(a, a2);
      b = IS_MIN ? fmin(b, b2) : fmax(b, b2);
    }
    double* red = sm + OFF_RED + red_sel * NW * RED;
    if (lane == 0) { red[warp * RED] = a; red[warp * RED + 1] = b; }
    __syncthreads();
#pragma unroll
    for (int wv = 0; wv < NW; ++wv) {
      a = IS_MIN ? fmin(a, red[wv * RED]) : fmax(a, red[wv * RED]);
      b = IS_MIN ? fmin(b, red[wv * RED + 1]) : fmax(b, red[wv * RED + 1]);
    }
    red_sel ^= 1;
  }

  // ---- factorisation -----------------------------------------------------------------------------------------
  __device__ __forceinline__ bool factorize() {
    // element-wise barrier weights
    const double dw0 = (has_w && valid) ? zw * (iw = 1.0 / w) : (iw = 1.0 / w, 0.0);
    double e = 0.0;
    ie = 1.0; ph = 0.0;
    if (has_u) {
      isp = 1.0 / sp; isq = 1.0 / sq;
      const double dp = zp * isp, dq = zq * isq;
      ie = 1.0 / (dp + dq + delta);
      ph = (dq - dp) * ie;
      e = (4.0 * dp * dq + 2.0 * delta * (dp + dq) + delta * delta) * ie;
    } else { isp = 1.0; isq = 1.0; }
    X(AD, k, i) = dw0 + delta; X(EE, k, i) = e; X(XR, k, i) = R; X(XPHI, k, i) = ph;
    __syncthreads();
    // conductance sweeps along the stages of one asset: left sweep by the stage-0 warps, right sweep by the
    // stage-(H-1) warps, concurrently (sequential in the stage index by nature)
    if (k == 0) {
      double hl = 0.0;
#pragma unroll
      for (int l = 0; l < H; ++l) {
        const double el = X(EE, l, i);
        double ql = 1.0, tl = 0.0;
        if (l > 0) { const double inv = 1.0 / (el + hl); ql = hl * inv; tl = el * inv; }
        X(QL, l, i) = ql; X(TL, l, i) = tl;
        hl = X(AD, l, i) + el * ql;
        X(HL, l, i) = hl;
      }
    }
    if (k == H - 1) {
      double hr = X(AD, H - 1, i);
      X(HR, H - 1, i) = hr;
      X(QR, 0, i) = 0.0; X(TR, 0, i) = 0.0;
#pragma unroll
      for (int l = H - 1; l >= 1; --l) {
        const double el = X(EE, l, i);
        const double inv = 1.0 / (el + hr);
        const double qr = hr * inv;
        X(QR, l, i) = qr; X(TR, l, i) = el * inv;
        hr = X(AD, l - 1, i) + el * qr;
        X(HR, l - 1, i) = hr;
      }
    }
    __syncthreads();
    {  // node / edge quantities of my own (k, i)
      const double gR = (k + 1 < H) ? X(EE, (k + 1 < H) ? k + 1 : 0, i) * X(QR, (k + 1 < H) ? k + 1 : 0, i) : 0.0;
      X(GJJ, k, i) = 1.0 / (X(HL, k, i) + gR);
      double fl = 1.0, fr = 0.0;
      const double hrk = X(HR, k, i);
      if (k > 0) { const double hlp = X(HL, k - 1, i); const double inv = 1.0 / (hlp + hrk); fl = hlp * inv; fr = hrk * inv; }
      X(FL, k, i) = fl; X(FR, k, i) = fr;
      X(VD, k, i) = 1.0 / (X(EE, k, i) + hrk * fl);
    }
    __syncthreads();
    // my rows of the Green's functions:  Gr[j] = G[k,j],  Dr[j] = D[k,j] (drop across my edge k, injection at j),
    // Dc[l] = D[l,k] (drop across edge l, injection at my node k),  DDr[j] = DD[k,j]
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const double gjj = X(GJJ, j, i);
      double g = gjj, gprev = gjj;               // gprev = G[k-1, j] when j < k
      if (j < k) {
#pragma unroll
        for (int mm = 1; mm < H; ++mm) if (mm > j && mm <= k) { gprev = g; g *= X(TR, mm, i); }
        Dr[j] = -gprev * X(QR, k, i);
      } else {
#pragma unroll
        for (int mm = H - 1; mm >= 1; --mm) if (mm <= j && mm > k) g *= X(TL, mm, i);
        Dr[j] = g * X(QL, k, i);
      }
      Gr[j] = g;
      // dipole across edge j, drop across my edge k
      const double V = X(VD, j, i);
      double dd;
      if (j == k) dd = V;
      else if (k > j) {
        double v = V * X(FL, j, i);
#pragma unroll
        for (int mm = 1; mm < H; ++mm) if (mm > j && mm < k) v *= X(TR, mm, i);
        dd = -v * X(QR, k, i);
      } else {
        double v = -V * X(FR, j, i);
#pragma unroll
        for (int mm = H - 1; mm >= 1; --mm) if (mm < j && mm > k) v *= X(TL, mm, i);
        dd = v * X(QL, k, i);
      }
      DDr[j] = dd;
    }
#pragma unroll
    for (int l = 0; l < H; ++l)       // D[l,k] from my own row of G (G is symmetric): l <= k: G[l,k] qL[l]; l > k: -G[l-1,k] qR[l]
      Dc[l] = (l <= k) ? Gr[l] * X(QL, l, i) : -Gr[(l >= 1) ? l - 1 : 0] * X(QR, l, i);
    // border matrix: thread (l = k, i) contributes the entries whose ROW stage is l, for every column stage j
    double c[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) c[q] = 0.0;
    if (valid) {
#pragma unroll
      for (int j = 0; j < H; ++j) {
        const double Rj = X(XR, j, i);
        c[j] = Rj * Gr[j];                                        // (1t_l, Rt_j)
        if (j <= k) { c[H + j] = R * Rj * Gr[j]; c[2 * H + j] = Gr[j]; }   // (Rt_l,Rt_j), (1t_l,1t_j)
        if (has_c) {
          const double dm = -ph * Dr[j];
          c[3 * H + j] = dm * Rj;                                  // (et_l, Rt_j)
          c[4 * H + j] = dm;                                       // (et_l, 1t_j)
          if (j <= k) c[5 * H + j] = ph * X(XPHI, j, i) * DDr[j] + ((j == k) ? ie : 0.0);   // (et_l, et_j)
        }
      }
    }
    const double tot = warp_transpose_reduce<6 * H>(c, lane);
    sm[OFF_KP + warp * 32 + lane] = tot;
    __syncthreads();
    if (warp - k * G == 0 && lane < 6 * H) {       // first warp of each stage combines its G partials and scatters
      double sacc = 0.0;
#pragma unroll
      for (int g2 = 0; g2 < G; ++g2) sacc += sm[OFF_KP + (k * G + g2) * 32 + lane];
      const int grp = lane / H, j = lane - grp * H;
      int r = -1, cc2 = -1;
      if (grp == 0) { r = H + k; cc2 = j; }
      else if (grp == 1 && j <= k) { r = k; cc2 = j; }
      else if (grp == 2 && j <= k) { r = H + k; cc2 = H + j; }
      else if (has_c && grp == 3) { r = 2 * H + k; cc2 = j; }
      else if (has_c && grp == 4) { r = 2 * H + k; cc2 = H + j; }
      else if (has_c && grp == 5 && j <= k) { r = 2 * H + k; cc2 = 2 * H + j; }
      if (r >= 0) {
        if (r == cc2) {                              // diagonal additions: 1/beta_k = rho_k^2, sc_k / zc_k
          if (grp == 1) sacc += rho * rho;
          if (grp == 5) sacc += sc / zc;
        }
        Kx(r, cc2) = sacc;
      }
    }
    __syncthreads();
    // Cholesky of the nb x nb border by warp 0 (lane = row); everyone else waits at the barrier below
    if (warp == 0) {
      for (int j = 0; j < nb; ++j) {
        const double djj = Kx(j, j);
        if (!(djj > 0.0)) { if (lane == 0) sm[OFF_K + NB * NB] = -1.0; break; }
        const double inv = 1.0 / sqrt(djj);
        if (lane > j && lane < nb) Kx(lane, j) *= inv;
        if (lane == j) { Kx(j, j) = djj * inv; sm[OFF_K + NB * NB + j] = inv; }
        __syncwarp();
        if (lane > j && lane < nb) {
          const double lij = Kx(lane, j);
          for (int k2 = j + 1; k2 <= lane; ++k2) Kx(lane, k2) -= lij * Kx(k2, j);
        }
        __syncwarp();
      }
    }
    __syncthreads();
    return sm[OFF_K + NB * NB] > 0.0;               // first inverse pivot is positive iff the factorisation went through
  }

  // t (in shared memory at OFF_T) <- K^{-1} t, by warp 0 (forward / backward substitution with the Cholesky factor;
  // an explicit inverse factor was tried and costs accuracy: the border's condition number reaches 1e9)
  __device__ __forceinline__ void k_solve_shared() {
    if (warp == 0) {
      double t = (lane < nb) ? sm[OFF_T + lane] : 0.0;
      for (int j = 0; j < nb; ++j) {
        const double yj = shfl_d(t, j) * sm[OFF_K + NB * NB + j];
        if (lane == j) t = yj;
        if (lane > j && lane < nb) t -= Kx(lane, j) * yj;
      }
      for (int j = nb - 1; j >= 0; --j) {
        const double xj = shfl_d(t, j) * sm[OFF_K + NB * NB + j];
        if (lane == j) t = xj;
        if (lane < j) t -= Kx(j, lane) * xj;
      }
      if (lane < NB) sm[OFF_T + lane] = (lane < nb) ? t : 0.0;
    }
  }

  // dw_k, dd_k of my (k, i) for the right-hand side currently in XGW / XPG (all stages of my asset)
  __device__ __forceinline__ void m0_rows(double& dwk, double& ddk) const {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const double gj = X(XGW, j, i), pj = X(XPG, j, i);
      a += Gr[j] * gj - Dc[j] * pj;
      b += Dr[j] * gj - DDr[j] * pj;
    }
    dwk = a; ddk = b;
  }

  // One Newton solve.  use_c: complementarity targets are held in dzw/dzp/dzq (registers) and cc (uniform).
  __device__ __forceinline__ void newton(bool use_c) {
    // right-hand side of my (k, i)
    const double cwv = (use_c && has_w) ? dzw : 0.0, cpv = (use_c && has_u) ? dzp : 0.0, cqv = (use_c && has_u) ? dzq : 0.0;
    double gwk = valid ? R * irho - nu : 0.0, guk = 0.0, tq = 0.0;
    if (valid) {
      if (has_w) gwk += cwv * iw;
      if (has_u) {
        const double a1 = cpv * isp, a2 = cqv * isq;
        tq = a1 - a2;
        guk = -lam + a1 + a2;
        if (has_c) guk -= cc * isc;
      }
    }
    X(XTQ, k, i) = tq;
    __syncthreads();
    gwk -= tq;
    if (k + 1 < H) gwk += X(XTQ, (k + 1 < H) ? k + 1 : 0, i);
    if (!valid) { gwk = 0.0; guk = 0.0; }
    X(XGW, k, i) = gwk; X(XPG, k, i) = ph * guk;
    __syncthreads();
    double dw0, dd0;
    m0_rows(dw0, dd0);
    {
      double v[3] = {valid ? R * dw0 : 0.0, valid ? dw0 : 0.0, (valid && has_c) ? guk * ie - ph * dd0 : 0.0};
      double S[3], T[3];
      reduce_sum<3>(v, S, T);
      if (warp - k * G == 0 && lane == 0) {
        sm[OFF_T + k] = S[0];
        sm[OFF_T + H + k] = S[1] + rp;            // t[H+k] = sum dw0 - q, q = -rp
        sm[OFF_T + 2 * H + k] = has_c ? S[2] : 0.0;
      }
    }
    __syncthreads();
    k_solve_shared();
    __syncthreads();
    const double yR = sm[OFF_T + k], yN = sm[OFF_T + H + k], yC = has_c ? sm[OFF_T + 2 * H + k] : 0.0;
    dnu = yN;
    if (has_c) { dsc = -yC * sc / zc; dzc = (cc * isc - zc) + yC; } else { dsc = 0.0; dzc = 0.0; }
    const double gw2 = valid ? gwk - (yR * R + yN) : 0.0;
    const double geff = valid ? guk - yC : 0.0;
    X(XGW, k, i) = gw2; X(XPG, k, i) = ph * geff;
    __syncthreads();
    double ddk;
    m0_rows(dw, ddk);
    if (valid) {
      dzw = has_w ? (cwv * iw - zw) - (zw * iw) * dw : 0.0;
      if (has_u) {
        const double dp = zp * isp, dq = zq * isq;
        dsp = (geff - (2.0 * dq + delta) * ddk) * ie;
        dsq = (geff + (2.0 * dp + delta) * ddk) * ie;
        dzp = (cpv * isp - zp) - dp * dsp;
        dzq = (cqv * isq - zq) - dq * dsq;
      } else { dsp = 0.0; dsq = 0.0; dzp = 0.0; dzq = 0.0; }
    } else { dw = 0.0; dsp = 0.0; dsq = 0.0; dzw = 0.0; dzp = 0.0; dzq = 0.0; }
    __syncthreads();                               // XGW / XPG are rewritten by the next solve
  }

  __device__ __forceinline__ void max_step(bool allow_short, double& ap, double& ad) {
    double p = 1.0, q = 1.0;
    auto lim = [](double v, double dv, double a0) {
      if (dv < 0.0 && fma(a0, dv, v) < 0.0) a0 = fmin(a0, -v / dv);
      return a0;
    };
    if (valid) {
      if (has_w) { p = lim(w, dw, p); q = lim(zw, dzw, q); }
      if (has_u) { p = lim(sp, dsp, p); p = lim(sq, dsq, p); q = lim(zp, dzp, q); q = lim(zq, dzq, q); }
    }
    if (has_c) { p = lim(sc, dsc, p); q = lim(zc, dzc, q); }
    if (allow_short) {
      double v[1] = {valid ? dw * R : 0.0}, S[1], T[1];
      reduce_sum<1>(v, S, T);
      p = lim(rho, S[0], p);
    }
    reduce_ext2<true>(p, q);
    ap = p; ad = q;
  }

  // Solve one problem.  R (gross return of my (k,i)) is set; w0s = sm + OFF_W0 holds the current weights of all
  // assets.  Returns the status; w holds my entry of the plan (tile(w_cur) on failure).
  __device__ __forceinline__ int solve(int N, double lam_, double tau_, bool allow_short, const IpmOptions& opt,
                                       int& iters, double (&kkt)[3]) {
    lam = lam_; tau = tau_; delta = opt.delta;
    has_u = (lam > 0.0) || (tau > 0.0);
    has_c = has_u && (tau > 0.0);
    has_w = !allow_short;
    nb = has_c ? 3 * H : 2 * H;
    iters = 0;
    kkt[0] = kkt[1] = kkt[2] = CUDART_NAN;
    const double w0 = valid ? sm[OFF_W0 + i] : 0.0;
    {
      double bad = (valid && !(isfinite(w0) && isfinite(R) && R > 0.0)) ? 1.0 : 0.0, zero = 0.0;
      reduce_ext2<false>(bad, zero);
      if (bad > 0.0) { w = w0; return ST_NONFINITE; }
    }
    // ---- initial point ------------------------------------------------------------------------------------------
    const double base = valid ? (allow_short ? w0 : fmax(w0, 0.0)) : 0.0;
    const double invN = 1.0 / (double)N;
    const double eps = (tau <= 0.0) ? 0.1 : fmin(0.1, tau / 8.0);
    double S1[1], T1[1];
    { double v[1] = {base}; reduce_sum<1>(v, S1, T1); }
    const double sb = S1[0];
    const double b = (sb > 0.0) ? base / sb : invN;
    w = valid ? (1.0 - eps) * b + eps * invN : 1.0;
    { double v[1] = {valid ? fabs(w - w0) : 0.0}; reduce_sum<1>(v, S1, T1); }
    const double absd0 = S1[0];
    sp = 1.0; sq = 1.0; sc = 1.0;
    if (has_u) {
      double dl0, dlk;
      if (tau > 0.0) {
        const double room0 = tau - absd0;
        if (!(room0 > 0.0)) { w = w0; kkt[0] = kkt[1] = kkt[2] = CUDART_INF; return ST_FAILED; }
        dl0 = room0 / (2.0 * N); dlk = tau / (2.0 * N);
      } else { dl0 = dlk = 0.05 * invN; }
      const double dk = (k == 0 && valid) ? w - w0 : 0.0;
      const double uk = (k == 0) ? fabs(dk) + dl0 : dlk;
      sp = uk - dk; sq = uk + dk;
      double v[1] = {valid ? uk : 0.0};
      reduce_sum<1>(v, S1, T1);
      sc = has_c ? tau - S1[0] : 1.0;
    }
    { double v[1] = {valid ? w * R : 0.0}; reduce_sum<1>(v, S1, T1); }
    rho = S1[0];
    const bool dual_start = has_w && (opt.dual_init > 0.0);
    if (dual_start) {
      const double zeta0 = has_c ? opt.dual_init : 0.0;
      // stage max of R / rho
      double mx = valid ? R / rho : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, shfl_xor_d(mx, o));
      double* red = sm + OFF_RED + red_sel * NW * RED;
      if (lane == 0) red[warp * RED] = mx;
      __syncthreads();
#pragma unroll
      for (int g2 = 0; g2 < G; ++g2) mx = fmax(mx, red[(k * G + g2) * RED]);
      red_sel ^= 1;
      nu = mx + opt.dual_init;
      zc = has_c ? zeta0 : 0.0;
      zw = valid ? (-R / rho + nu) : 0.0;
      zp = has_u ? 0.5 * fmax(lam + zeta0, opt.dual_init) : 0.0;   // floor: see oracle
      zq = zp;
    } else {
      nu = 1.0; zc = has_c ? opt.mu0 / sc : 0.0;
      zw = (has_w && valid) ? opt.mu0 / w : 0.0;
      zp = has_u ? opt.mu0 / sp : 0.0;
      zq = has_u ? opt.mu0 / sq : 0.0;
    }
    const double mcount = (has_w ? (double)H * N : 0.0) + (has_u ? 2.0 * H * N : 0.0) + (has_c ? (double)H : 0.0);
    int status = ST_FAILED;
    for (int it = 1; it <= opt.max_iter + 1; ++it) {
      iters = it;
      // ---- residuals ------------------------------------------------------------------------------------------------
      double gapc = 0.0;
      if (valid) {
        if (has_w) gapc += w * zw;
        if (has_u) gapc += sp * zp + sq * zq;
      }
      if (has_c && warp - k * G == 0 && lane == 0) gapc += sc * zc;
      double v[3] = {valid ? w * R : 0.0, valid ? w : 0.0, gapc}, S[3], T[3];
      reduce_sum<3>(v, S, T);
      rho = S[0]; irho = 1.0 / rho; isc = has_c ? 1.0 / sc : 0.0;
      rp = S[1] - 1.0;
      const double gap = T[2];
      // dual residual of my (k, i): needs (zp - zq) of the next stage
      X(XTQ, k, i) = zp - zq;
      __syncthreads();
      double dres = 0.0, pres = fabs(rp);
      if (valid) {
        const double yn = (k + 1 < H) ? X(XTQ, (k + 1 < H) ? k + 1 : 0, i) : 0.0;
        dres = fabs(-R * irho - zw + (zp - zq) - yn + nu);
        if (has_u) dres = fmax(dres, fabs(lam - zp - zq + (has_c ? zc : 0.0)));
      }
      reduce_ext2<false>(pres, dres);
      kkt[0] = pres; kkt[1] = dres; kkt[2] = gap;
      if (!isfinite(dres + gap)) break;
      if (pres < opt.tol && dres < opt.tol_dual && gap < opt.tol) { status = ST_OPTIMAL; break; }
      if (pres < opt.tol && gap < 1e-6 * opt.tol && dres < 1e-6) { status = ST_INACCURATE; break; }
      if (it == opt.max_iter + 1) break;
      const double mu = gap / fmax(mcount, 1.0);
      if (pres < opt.tol && gap < opt.tol) delta = fmax(0.3 * delta, 1e-9);
      if (!factorize()) break;
      cc = 0.0;
#pragma unroll 1
      for (int phase = (mcount > 0.0 ? 0 : 1); phase < 2; ++phase) {
        const bool use_c = (phase == 1) && (mcount > 0.0);
        newton(use_c);
        double aa = 1.0, ab = 1.0;
        if (mcount > 0.0 || allow_short) max_step(allow_short, aa, ab);
        if (phase == 0) {
          double g2 = 0.0;
          if (valid) {
            if (has_w) g2 += (w + aa * dw) * (zw + ab * dzw);
            if (has_u) g2 += (sp + aa * dsp) * (zp + ab * dzp) + (sq + aa * dsq) * (zq + ab * dzq);
          }
          if (has_c && warp - k * G == 0 && lane == 0) g2 += (sc + aa * dsc) * (zc + ab * dzc);
          double vv[1] = {g2}, SS[1], TT[1];
          reduce_sum<1>(vv, SS, TT);
          const double ratio = (gap > 0.0) ? fmin(1.0, fmax(TT[0] / gap, 0.0)) : 0.0;
          const double smu = ratio * ratio * ratio * mu;
          const double dmp = fmin(1.0, fmin(aa, ab) * (1.0 / kCorrFull));
          dzw = has_w ? smu - dmp * dw * dzw : 0.0;           // complementarity targets replace the affine dual steps
          dzp = has_u ? smu - dmp * dsp * dzp : 0.0;
          dzq = has_u ? smu - dmp * dsq * dzq : 0.0;
          cc = has_c ? smu - dmp * dsc * dzc : 0.0;
        } else {
          const bool stepped = (mcount > 0.0 || allow_short);
          const double pa = stepped ? fmin(1.0, opt.step_frac * aa) : 1.0;
          const double pb = stepped ? fmin(1.0, opt.step_frac * ab) : 1.0;
          if (valid) {
            w += pa * dw;
            if (has_w) zw += pb * dzw;
            if (has_u) { sp += pa * dsp; sq += pa * dsq; zp += pb * dzp; zq += pb * dzq; }
          }
          nu += pb * dnu;
          if (has_c) { sc += pa * dsc; zc += pb * dzc; }
        }
      }
    }
    if (status != ST_OPTIMAL && isfinite(kkt[1] + kkt[2]) && kkt[0] < kLoosePres && kkt[1] < kLooseDres && kkt[2] < kLooseGap)
      status = ST_INACCURATE;
    if (status == ST_FAILED) w = w0;
    return status;
  }

  // maximised objective (mpc.py:104) of the plan held in the w registers; same value in every thread
  __device__ __forceinline__ double objective() {
    const double w0 = valid ? sm[OFF_W0 + i] : 0.0;
    __syncthreads();
    X(XGW, k, i) = w;
    __syncthreads();
    const double prev = (k == 0) ? w0 : X(XGW, (k == 0) ? 0 : k - 1, i);
    double v[2] = {valid ? w * R : 0.0, valid ? fabs(w - prev) : 0.0}, S[2], T[2];
    reduce_sum<2>(v, S, T);
    // log(rho_k) summed over stages: lane 0 of the first warp of each stage contributes
    double lv[1] = {(warp - k * G == 0 && lane == 0) ? log(S[0]) : 0.0}, LS[1], LT[1];
    reduce_sum<1>(lv, LS, LT);
    return LT[0] - lam * T[1];
  }
};

}  // namespace kmpc
