// Lane-per-asset layout of the fp64 interior-point MPC solver (same central path and Newton system as
// oracle/mpc_oracle.py::solve_structured(apply="sweep"); see mpc_common.cuh for the program).
//
// Why a third layout (profiles/r1_backtest_cta_kernel.txt): the warp- and CTA-per-problem kernels execute
// ~350 k warp instructions per decision of which 17 % are fp64 math; the rest is predication on a runtime stage
// index, O(H^2) Green's-function columns rebuilt for every matrix-vector product, full-precision divisions, a
// Cholesky that makes ~105 dependent shared-memory round trips, and block-wide barriers (48 % of all stalls).
//
//   thread i  = asset i; block = G warps = one problem (N <= 32 G).  Every loop over the H stages is unrolled:
//               the iterate (R, w, sp, sq, zw, zp, zq), the element-wise barrier factors and the search direction
//               of all stages of one asset are statically indexed REGISTERS of its thread; H independent
//               dependency chains per thread give the ILP that the low occupancy needs.
//   M0^{-1}   = the per-asset SPD tridiagonal (path network) is applied in O(H) by two Norton-equivalent sweeps
//               (left sources JL, right sources JR) and a two-sided split at every edge: node potentials AND the
//               drops across the edges come out of products of factors in [0,1] without differencing potentials
//               (oracle: _path_sweep; as accurate as the explicit Green's functions, 9 H flops instead of ~7 H^2).
//   border K  = the <= 3H x 3H Schur complement needs the explicit entries sum_i c_i G_i[l,j]: every thread
//               emits its <= 3H^2 + 3H(H+1)/2 products into a [entry][thread] shared-memory tile (conflict free),
//               lane e of every warp adds up row e over its 32 columns (16 LDS.128 + 32 DADD for 32 entries; a
//               shuffle butterfly costs ~8 instructions per entry), warp 0 combines the G partials.
//   L D L'    = warp 0, lane = row, the row lives in registers, columns are exchanged by shuffles: 15 dependent
//               steps instead of 105 dependent shared-memory round trips, no square root; the factor goes back to
//               shared memory for the two triangular solves of each Newton system.
//   division  = MUFU.RCP64H + one third-order Newton step (4 instructions, <= 2 ulp) for the ~40 reciprocals
//               per asset and iteration; the step-length ratio tests multiply by reciprocals and keep a running
//               integer maximum of the high words (no divergent division, no DSETP/FSEL chains).
//   flags     = FIX = true instantiates the solver for the reference defaults (long-only, lam > 0, tau > 0) with
//               the structure flags as compile-time constants; LOC = true keeps factors / targets thread-private
//               (H = 10 or N > 128, see below).
//   pieces    = begin / check / factor_a / factor_b / newton_phase, so that the persistent backtest kernel
//               (mpc_lane_kernels.cuh) can walk several problems of one SM through an iteration together.
#pragma once
#include "mpc_common.cuh"

// tuning switches (scripts/ab_bench.py builds variants with -D flags; the defaults are the product)
#ifndef KMPC_BW_SPREAD
#define KMPC_BW_SPREAD 1        // border warp chosen per slot so that the four chains of an SM use four sub-partitions
#endif
#ifndef KMPC_RCP_QUADRATIC
#define KMPC_RCP_QUADRATIC 0    // experiment: one quadratic Newton step after MUFU.RCP64H (~1e-12 relative) instead of the cubic one
#endif

namespace kmpc {

// Block-uniform conditions are evaluated from shared-memory values; the vote makes them provably warp-uniform so
// that the shuffles downstream compile without divergence handling (WARPSYNC / ENDCOLLECTIVE per shuffle).
__device__ __forceinline__ bool uni(bool c) { return __any_sync(kFull, c) != 0; }

__device__ __forceinline__ double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));   // ~20 correct bits
  const double e = fma(-x, y, 1.0);
#if KMPC_RCP_QUADRATIC
  return fma(y, e, y);
#else
  return fma(y, fma(e, e, e), y);                            // relative error ~e^3
#endif
}
// a / b to within ~1.5 ulp in 7 instructions.  Used where a decision is STARTED or BOOKED: that code runs once per
// decision, each problem of a block at its own time, so it is always fetched cold (the Newton loop has evicted it)
// and its size, not its latency, is what it costs; an IEEE division is ~20 inline instructions plus a slow-path call.
__device__ __forceinline__ double div_fast(double a, double b) { return a * rcp_fast(b); }

// Position r * NB + c (lower triangle) of the K entries in the order LaneIpm::assemble_K emits them.
template <int H>
struct KMap {
  static constexpr int NB = 3 * H;
  static constexpr int N1 = H * H + H * (H + 1);                 // entries without the cap rows
  static constexpr int NK = 3 * H * H + 3 * H * (H + 1) / 2;
  int rc[NK];
  constexpr KMap() : rc() {
    int e = 0;
    for (int j = 0; j < H; ++j)
      for (int l = j; l < H; ++l) {
        rc[e++] = (H + l) * NB + j;                              // (1t_l, Rt_j)
        if (l != j) rc[e++] = (H + j) * NB + l;                  // (1t_j, Rt_l)
        rc[e++] = l * NB + j;                                    // (Rt_l, Rt_j)
        rc[e++] = (H + l) * NB + H + j;                          // (1t_l, 1t_j)
      }
    for (int l = 0; l < H; ++l)
      for (int j = 0; j < H; ++j) {
        rc[e++] = (2 * H + l) * NB + j;                          // (et_l, Rt_j)
        rc[e++] = (2 * H + l) * NB + H + j;                      // (et_l, 1t_j)
      }
    for (int j = 0; j < H; ++j)
      for (int l = j; l < H; ++l) rc[e++] = (2 * H + l) * NB + 2 * H + j;   // (et_l, et_j)
  }
};
template <int H>
__device__ __constant__ KMap<H> g_kmap{};

// Tensor-core assembly of the border matrix (H <= 5, non-LOC kernels).  Per asset the Green's functions of the path
// network are semi-separable: with T_k = tR_1 ... tR_k,  G[l][j] = T_l * (gjj_j / T_j) for l >= j, and the edge drops
// D, DD factor the same way (all products / quotients of positive numbers: componentwise accurate).  Every entry of K is
// therefore ONE inner product over the assets of a left column X[.][m] with a right column Y[.][n]:
//   X = [ P_R(H) | P_1(H) | Pe2(H) ],   P_R[k] = R_k T_k,  P_1[k] = T_k,  Pe2[k] = ph_k qR_k T_{k-1}  (0 for k = 0)
//   Y = [ Q_R(H) | Q_1(H) | Qe1(H) | Qe3(H-1) | d(H) ],  Q_R[k] = R_k gjj_k / T_k,  Q_1[k] = gjj_k / T_k,
//       Qe1[k] = -ph_k qL_k gjj_k / T_k,  Qe3[k] = -ph_k vd_k fL_k / T_k,  d[k] = ph_k^2 vd_k + ie_k (diagonal of the
//       cap block; summed against the constant column P_1[0] = 1).
// K = X' Y is a [3H x 5H-1] product with the asset axis as the contraction: fp64 mma.sync.m8n8k4 (DMMA), 2 x 3 tiles of
// 8 x 8 at H = 5, each warp contracting over its own 32 assets.  KDst maps a position of that product to its entry of
// the lower triangle of K (or -1: the 60 % of the products that the triangle structure does not need).
template <int H>
struct KDst {
  static constexpr int NB = 3 * H;
  static constexpr int MT = (3 * H + 7) / 8;            // 8-row tiles of X' (M side)
  static constexpr int NTL = (5 * H - 1 + 7) / 8;       // 8-column tiles of Y (N side)
  short dst[MT * NTL * 2][32];                          // [tile * 2 + element][lane]
  static constexpr int of(int m, int n) {
    if (m >= 3 * H || n >= 5 * H - 1) return -1;
    const int a = m / H, l = m % H;                     // a: 0 P_R, 1 P_1, 2 Pe2
    if (n >= 4 * H - 1) {                               // diagonal slots, summed against the ones column P_1[0]
      const int k = n - (4 * H - 1);
      return (a == 1 && l == 0) ? (2 * H + k) * NB + 2 * H + k : -1;
    }
    const int b = n / H, j = n % H;                     // b: 0 Q_R, 1 Q_1, 2 Qe1, 3 Qe3
    if (a <= 1 && b <= 1) {
      if (l < j) return -1;
      if (a == 0 && b == 0) return l * NB + j;                          // (Rt_l, Rt_j)
      if (a == 1 && b == 0) return (H + l) * NB + j;                    // (1t_l, Rt_j)
      if (a == 0 && b == 1) return (l > j) ? (H + j) * NB + l : -1;     // (1t_j, Rt_l)
      return (H + l) * NB + H + j;                                      // (1t_l, 1t_j)
    }
    if (a == 2) {                                       // edge l, source node / edge j strictly to its left
      if (!(l > j)) return -1;
      if (b == 0) return (2 * H + l) * NB + j;                          // (et_l, Rt_j)
      if (b == 1) return (2 * H + l) * NB + H + j;                      // (et_l, 1t_j)
      if (b == 3) return (2 * H + l) * NB + 2 * H + j;                  // (et_l, et_j)
      return -1;
    }
    if (b == 2) return (j <= l) ? (2 * H + j) * NB + (a == 0 ? l : H + l) : -1;   // edge j at or left of node l
    return -1;
  }
  constexpr KDst() : dst() {
    for (int mt = 0; mt < MT; ++mt)
      for (int nt = 0; nt < NTL; ++nt)
        for (int e = 0; e < 2; ++e)
          for (int lane = 0; lane < 32; ++lane)
            dst[(mt * NTL + nt) * 2 + e][lane] = (short)of(8 * mt + lane / 4, 8 * nt + 2 * (lane % 4) + e);
  }
};
template <int H>
__device__ const KDst<H> g_kdst{};

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// LOC = true ("large" problems, e.g. 500 assets x 10 stages): the per-asset sweep factors and corrector targets are
// thread-private arrays instead of shared-memory columns (a problem's 100+ rows x 512 assets do not fit an SM's
// shared memory); together with the iterate they exceed the register file and ptxas keeps the excess in local
// memory (L1/L2-resident).  Same code, same arithmetic; slower per decision, only the reduction tile, K and the
// stage scalars stay in shared memory.
template <int H, int G, bool LOC = false, bool FIX = false>
struct LaneIpm {
  static constexpr int NT = 32 * G;                 // threads per problem
  static constexpr int NB = 3 * H;
  static constexpr int LD = NT + 2;                 // row stride of the reduction tile: 16-byte rows, LD/2 odd
  static constexpr int KB = (NB + 1 > 30) ? 32 : 30;   // entries per K-assembly batch (<= 32)
  static constexpr int NK = KMap<H>::NK, N1 = KMap<H>::N1;
  static_assert(NB + 1 <= 32, "H too large: the border must fit one row per lane");

  // ---- shared memory of one problem (doubles) -------------------------------------------------------------------
  // sweep factors [stage][NT]; stage 0 of QL,TL,QR,TR,FL,FR is constant (1,0,0,0,1,0) and not stored
  enum : int { F_GJJ, F_VD, F_QL, F_TL, F_QR, F_TR, F_FL, F_FR, NFAC };
  static constexpr int FAC_ROWS = 2 * H + 6 * (H - 1);
  enum : int { T_CW, T_CP, T_CQ, NTGT };                                    // complementarity targets [H][NT]
  enum : int { U_NU, U_SC, U_ZC, U_RHO, U_IRHO, U_ISC, U_RP, U_CC, NUNI };
  static constexpr int OFF_FAC = 0;
  static constexpr int OFF_TILE = OFF_FAC + (LOC ? 0 : FAC_ROWS * NT);      // reduction tile, rows of LD doubles
  static constexpr int SMALL_ROWS = NB + 1;                                 // rows usable while the targets are live
  // border assembly on the fp64 tensor cores (see KDst) where the border fits one 16 x 24 product; else the tile path
  static constexpr bool DMMA = !LOC && H <= 5;
  static constexpr int MT = KDst<(H <= 5 ? H : 1)>::MT, NTL = KDst<(H <= 5 ? H : 1)>::NTL;
  static constexpr int LDX = NT + 4;                  // row stride of the X|Y columns: LDX mod 16 = 4, conflict-free LDS.64
  static constexpr int XY_DOUBLES = 8 * (MT + NTL) * LDX;
  static constexpr int PART_DOUBLES = (G - 1) * MT * NTL * 2 * 32;   // partial products of warps 1..G-1
  static constexpr int TILE_A = DMMA ? XY_DOUBLES + PART_DOUBLES : KB * LD;
  static constexpr int TGT_DOUBLES = LOC ? 0 : NTGT * H * NT;
  static constexpr int TILE_B = SMALL_ROWS * LD + TGT_DOUBLES;
  static constexpr int TILE_DOUBLES = ((TILE_A > TILE_B ? TILE_A : TILE_B) + 1) & ~1;
  static constexpr int OFF_TGT = OFF_TILE + TILE_DOUBLES - TGT_DOUBLES;    // targets = tail of the tile
  static constexpr int OFF_K = OFF_TILE + TILE_DOUBLES;                     // [NB*NB] K / unit-lower factor L, [NB] 1/D
  static constexpr int OFF_T = OFF_K + NB * NB + NB + ((NB * NB + NB) & 1); // [32] border right-hand side / solution
  static constexpr int OFF_P = OFF_T + 32;                                  // [2][G][32] per-warp partials
  static constexpr int OFF_U = OFF_P + 2 * G * 32;                          // [NUNI][H] stage scalars
  static constexpr int OFF_FLAG = OFF_U + NUNI * H + ((NUNI * H) & 1);      // [2]
  static constexpr int SMEM_DOUBLES = OFF_FLAG + 2;

  // ---- registers of thread i (asset i) ----------------------------------------------------------------------------
  double R[H], w[H], sp[H], sq[H], zw[H], zp[H], zq[H];      // iterate
  double iw[H], isp[H], isq[H], ie[H], ph[H];                // element-wise factors of the current iterate
  mutable double fac_[LOC ? FAC_ROWS : 1], tgt_[LOC ? NTGT * H : 1];   // LOC only: thread-private factors / targets
  double* sm;
  int tid, lane, warp, psel, bar_id;
  int bw;                             // the warp of this problem that factorises and solves the border system
  bool valid, has_w_, has_u_, has_c_, allow_short_, fact_ok_;
  double lam, tau, delta;
  // state of the solve in progress (begin / check / factor_a / factor_b / newton_phase)
  double w0_, mu_, gap_, mcount_, kkt_[3];
  int it_, it0_, nretry_;
  bool robust_;                       // second attempt (see kRobust* in mpc_common.cuh)

  // `slot`: which of the problems that share this thread block (its threads are [slot*NT, (slot+1)*NT), its
  // named barrier is 1 + slot; barrier 0 stays free for block-wide lockstep points of the caller).
#ifdef KMPC_LANE_PROFILE
  // phase clocks of thread 0 of the block (tuning builds only; see profiles/README.md)
  long long* prof_ = nullptr; long long tl_ = 0;
#define KMPC_PROF(S, i) if ((S).prof_) { const long long n_ = clock64(); (S).prof_[i] += n_ - (S).tl_; (S).tl_ = n_; }
#else
#define KMPC_PROF(S, i)
#endif
  __device__ __forceinline__ void bind(double* smem_slice, int n_assets, int slot) {
    sm = smem_slice; tid = (int)threadIdx.x - slot * NT; lane = tid & 31; psel = 0; bar_id = 1 + slot;
    warp = __shfl_sync(kFull, tid >> 5, 0);        // provably warp-uniform: branches on it need no reconvergence code
    // The border work (L D L', triangular solves) is one warp's dependent chain.  Warp w of the block issues on SM
    // sub-partition w % 4: with "warp 0 of every slot" the chains of the four slots of an SM share two sub-partitions;
    // picking the border warp by slot gives every chain a sub-partition of its own.
    bw = (KMPC_BW_SPREAD && DMMA && G > 1 && G <= 4) ? __shfl_sync(kFull, (slot / (4 / (G <= 4 ? G : 4))) % G, 0) : 0;
    valid = tid < n_assets;
    fact_ok_ = true; it_ = 0; it0_ = 0; robust_ = false;
  }
  __device__ __forceinline__ double& FAC(int arr, int k) const {      // arr >= F_QL requires k >= 1
    const int row = (arr < F_QL) ? arr * H + k : 2 * H + (arr - F_QL) * (H - 1) + (k - 1);
    if (LOC) return fac_[LOC ? row : 0];
    return sm[OFF_FAC + row * NT + tid];
  }
  __device__ __forceinline__ double& TGT(int arr, int k) const {
    if (LOC) return tgt_[LOC ? arr * H + k : 0];
    return sm[OFF_TGT + (arr * H + k) * NT + tid];
  }
  // FIX: the caller guarantees lam > 0, tau > 0 and long-only weights (the reference defaults, mpc.py:17-25): the three
  // structure flags are compile-time constants and every select / branch on them disappears.
  __device__ __forceinline__ bool hw() const { return FIX ? true : has_w_; }
  __device__ __forceinline__ bool hu() const { return FIX ? true : has_u_; }
  __device__ __forceinline__ bool hc() const { return FIX ? true : has_c_; }
  __device__ __forceinline__ bool ash() const { return FIX ? false : allow_short_; }
  __device__ __forceinline__ double& U(int arr, int k) const { return sm[OFF_U + arr * H + k]; }
  __device__ __forceinline__ void sync() const {
    if (G == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NT) : "memory");
  }

  // ---- reductions over the threads of the problem --------------------------------------------------------------
  // Sum of row `lane` of the tile over this warp's 32 columns.
  __device__ __forceinline__ double row_partial(int rows) const {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (lane < rows) {
      const double2* p = reinterpret_cast<const double2*>(sm + OFF_TILE + lane * LD + 32 * warp);
#pragma unroll
      for (int m = 0; m < 16; m += 2) {
        const double2 a = p[m], b = p[m + 1];
        s0 += a.x; s1 += a.y; s2 += b.x; s3 += b.y;
      }
    }
    return (s0 + s1) + (s2 + s3);
  }
  // NV per-thread values (0 in padding threads) -> per-warp partial sums in P; ptotal(e) then gives the total of
  // entry e (same bits in every thread) until the next tile_reduce.
  template <int NV>
  __device__ __forceinline__ void tile_reduce(const double (&v)[NV]) {
    static_assert(NV <= SMALL_ROWS, "tile too small");
#pragma unroll
    for (int e = 0; e < NV; ++e) sm[OFF_TILE + e * LD + tid] = v[e];
    sync();
    const double s = row_partial(NV);
    psel ^= 1;
    sm[OFF_P + psel * G * 32 + warp * 32 + lane] = s;
    sync();
  }
  __device__ __forceinline__ double ptotal(int e) const {
    const double* P = sm + OFF_P + psel * G * 32;
    double t = P[e];
#pragma unroll
    for (int g = 1; g < G; ++g) t += P[g * 32 + e];
    return t;
  }
  // rows [0, NMAX) are max-reduced (non-negative values, 0 in padding threads), rows [NMAX, NV) summed
  template <int NMAX, int NV>
  __device__ __forceinline__ void tile_reduce_mixed(const double (&v)[NV]) {
    static_assert(NV <= SMALL_ROWS, "tile too small");
#pragma unroll
    for (int e = 0; e < NV; ++e) sm[OFF_TILE + e * LD + tid] = v[e];
    sync();
    double s = 0.0;
    if (lane < NV) {
      const double2* p = reinterpret_cast<const double2*>(sm + OFF_TILE + lane * LD + 32 * warp);
      double m0 = 0.0, m1 = 0.0, s0 = 0.0, s1 = 0.0;
#pragma unroll 2                                     // once per decision: keep the code small (see div_fast)
      for (int m = 0; m < 16; ++m) {
        const double2 a = p[m];
        m0 = fmax(m0, a.x); m1 = fmax(m1, a.y); s0 += a.x; s1 += a.y;
      }
      s = (lane < NMAX) ? fmax(m0, m1) : s0 + s1;
    }
    psel ^= 1;
    sm[OFF_P + psel * G * 32 + warp * 32 + lane] = s;
    sync();
  }
  __device__ __forceinline__ double ptotal_max(int e) const {
    const double* P = sm + OFF_P + psel * G * 32;
    double t = P[e];
#pragma unroll
    for (int g = 1; g < G; ++g) t = fmax(t, P[g * 32 + e]);
    return t;
  }
  template <int NV>
  __device__ __forceinline__ void block_sum(const double (&v)[NV], double (&out)[NV]) {
    tile_reduce<NV>(v);
#pragma unroll
    for (int e = 0; e < NV; ++e) out[e] = ptotal(e);
  }
  // Maxima of non-negative quantities are taken on the HIGH WORDS of the doubles as integers (one VIMNMX instead of
  // DSETP + 2 FSEL; a negative double has a negative high word and drops out against the initial 0) and rounded UP
  // when converted back: the result is >= the true maximum by at most 2^-20 relative, which errs on the safe side
  // for a step-length ratio and for a residual that is compared with a tolerance.
  static __device__ __forceinline__ int hi_of(double x) { return __double2hiint(x); }
  static __device__ __forceinline__ double hi_up(int h) { return __hiloint2double(h + 1, 0); }
  __device__ __forceinline__ void block_max2i(int& a, int& b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = max(a, __shfl_xor_sync(kFull, a, o));
      b = max(b, __shfl_xor_sync(kFull, b, o));
    }
    if (G > 1) {
      psel ^= 1;
      int* P = reinterpret_cast<int*>(sm + OFF_P + psel * G * 32);
      if (lane == 0) { P[warp * 2] = a; P[warp * 2 + 1] = b; }
      sync();
#pragma unroll
      for (int g = 0; g < G; ++g) { a = max(a, P[g * 2]); b = max(b, P[g * 2 + 1]); }
    }
  }
  __device__ __forceinline__ double block_sum1(double a) {
    a = warp_sum(a);
    if (G > 1) {
      psel ^= 1;
      double* P = sm + OFF_P + psel * G * 32;
      if (lane == 0) P[warp * 32] = a;
      sync();
      a = P[0];
#pragma unroll
      for (int g = 1; g < G; ++g) a += P[g * 32];
    }
    return a;
  }

  // ---- M0^{-1}: two-sided sweeps (oracle: _path_sweep) ---------------------------------------------------------
  __device__ __forceinline__ void m0_apply(const double (&gw)[H], const double (&pg)[H], double (&dw)[H],
                                           double (&dd)[H]) const {
    double JL[H], JR[H], inc[H];
    JL[0] = gw[0] - pg[0];
#pragma unroll
    for (int k = 1; k < H; ++k) JL[k] = fma(FAC(F_TL, k), JL[k - 1], fma(-FAC(F_QL, k), pg[k], gw[k]));
    JR[H - 1] = gw[H - 1]; inc[H - 1] = 0.0;
#pragma unroll
    for (int k = H - 1; k >= 1; --k) {
      inc[k - 1] = fma(FAC(F_TR, k), JR[k], FAC(F_QR, k) * pg[k]);
      JR[k - 1] = gw[k - 1] + inc[k - 1];
    }
#pragma unroll
    for (int k = 0; k < H; ++k) {
      dw[k] = FAC(F_GJJ, k) * (JL[k] + inc[k]);
      double t = JR[k] - pg[k];                                        // fL[0] = 1, fR[0] = 0
      if (k > 0) t = fma(-FAC(F_FR, k), JL[(k > 0) ? k - 1 : 0], fma(FAC(F_FL, k), JR[k], -pg[k]));
      dd[k] = FAC(F_VD, k) * t;
    }
  }

  // ---- factorisation ---------------------------------------------------------------------------------------------
  // One batch of the K assembly is complete in the tile: add up, combine the warps, scatter into K.
  __device__ __forceinline__ void flush_batch(int first, int count) {
    sync();
    double s = row_partial(count);
    if (G > 1) {
      psel ^= 1;
      double* P = sm + OFF_P + psel * G * 32;
      if (warp > 0) P[warp * 32 + lane] = s;
      sync();
      if (warp == 0) {
#pragma unroll
        for (int g = 1; g < G; ++g) s += P[g * 32 + lane];
      }
    } else {
      __syncwarp();
    }
    if (warp == 0 && lane < count) sm[OFF_K + g_kmap<H>.rc[first + lane]] = s;
  }

  // element-wise barrier factors, conductance sweeps, border matrix K (shared memory)
  __device__ __forceinline__ void factor_a() {
    double ad[H], e[H];
#pragma unroll
    for (int k = 0; k < H; ++k) {
      iw[k] = rcp_fast(w[k]);
      ad[k] = (hw() ? zw[k] * iw[k] : 0.0) + delta;
      if (hu()) {
        isp[k] = rcp_fast(sp[k]); isq[k] = rcp_fast(sq[k]);
        const double dp = zp[k] * isp[k], dq = zq[k] * isq[k];
        const double s = dp + dq;
        ie[k] = rcp_fast(s + delta);
        ph[k] = (dq - dp) * ie[k];
        e[k] = fma(4.0 * dp, dq, fma(2.0 * delta, s, delta * delta)) * ie[k];
      } else { isp[k] = 1.0; isq[k] = 1.0; ie[k] = 1.0; ph[k] = 0.0; e[k] = 0.0; }
    }
    // conductance sweeps (per thread, sequential in the stage index by nature)
    double hL[H], hR[H], qL[H], tL[H], qR[H], tR[H], gjj[H], vd[H], fL[H], fR[H];
    qL[0] = 1.0; tL[0] = 0.0; hL[0] = ad[0] + e[0];
#pragma unroll
    for (int l = 1; l < H; ++l) {
      const double inv = rcp_fast(e[l] + hL[l - 1]);
      qL[l] = hL[l - 1] * inv; tL[l] = e[l] * inv;
      hL[l] = fma(e[l], qL[l], ad[l]);
    }
    hR[H - 1] = ad[H - 1]; qR[0] = 0.0; tR[0] = 0.0;
#pragma unroll
    for (int l = H - 1; l >= 1; --l) {
      const double inv = rcp_fast(e[l] + hR[l]);
      qR[l] = hR[l] * inv; tR[l] = e[l] * inv;
      hR[l - 1] = fma(e[l], qR[l], ad[l - 1]);
    }
#pragma unroll
    for (int k = 0; k < H; ++k) {
      const double gR = (k + 1 < H) ? e[(k + 1 < H) ? k + 1 : 0] * qR[(k + 1 < H) ? k + 1 : 0] : 0.0;
      gjj[k] = rcp_fast(hL[k] + gR);
      fL[k] = 1.0; fR[k] = 0.0;
      if (k > 0) {
        const double inv = rcp_fast(hL[(k > 0) ? k - 1 : 0] + hR[k]);
        fL[k] = hL[(k > 0) ? k - 1 : 0] * inv; fR[k] = hR[k] * inv;
      }
      vd[k] = rcp_fast(fma(hR[k], fL[k], e[k]));
      if (!valid) { gjj[k] = 0.0; vd[k] = 0.0; ie[k] = 0.0; }     // padding lanes contribute exact zeros everywhere
      FAC(F_GJJ, k) = gjj[k]; FAC(F_VD, k) = vd[k];
      if (k > 0) {
        FAC(F_QL, k) = qL[k]; FAC(F_TL, k) = tL[k]; FAC(F_QR, k) = qR[k]; FAC(F_TR, k) = tR[k];
        FAC(F_FL, k) = fL[k]; FAC(F_FR, k) = fR[k];
      }
    }
    KMPC_PROF(*this, 2)
    if constexpr (DMMA) {
      assemble_K_dmma(gjj, vd, fL, qL, qR, tR);
    } else {
    // ---- border matrix: emit the entries in KMap order, KB at a time (H = 10 / thread-private factors) ----------
    if (warp == 0) {                                 // clear K (rows >= 2H stay identity when there is no cap)
      for (int q = lane; q < NB * NB; q += 32) sm[OFF_K + q] = (q / NB == q % NB) ? 1.0 : 0.0;
    }
    double* tile = sm + OFF_TILE + tid;
    int en = 0;                                      // compile-time after unrolling
    auto emit = [&](double val) {
      tile[(en % KB) * LD] = val;
      ++en;
    };
#define KMPC_FLUSH_IF_FULL()                                             \
    if (en % KB == 0) flush_batch(en - KB, KB);
    double Gm[H][H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
      double g = gjj[j];
#pragma unroll
      for (int l = j; l < H; ++l) {
        if (l > j) g *= tR[l];
        Gm[l][j] = g;
        emit(R[j] * g); KMPC_FLUSH_IF_FULL();
        if (l != j) { emit(R[l] * g); KMPC_FLUSH_IF_FULL(); }
        emit(R[l] * R[j] * g); KMPC_FLUSH_IF_FULL();
        emit(g); KMPC_FLUSH_IF_FULL();
      }
    }
    if (hc()) {
#pragma unroll
      for (int l = 0; l < H; ++l) {
#pragma unroll
        for (int j = 0; j < H; ++j) {
          // D[l][j]: drop across edge l for a unit injection at node j
          const double D = (l <= j) ? Gm[j][l] * qL[l] : -Gm[(l > 0) ? l - 1 : 0][j] * qR[l];
          const double dm = -ph[l] * D;
          emit(dm * R[j]); KMPC_FLUSH_IF_FULL();
          emit(dm); KMPC_FLUSH_IF_FULL();
        }
      }
#pragma unroll
      for (int j = 0; j < H; ++j) {
        double v = vd[j] * fL[j];
#pragma unroll
        for (int l = j; l < H; ++l) {
          double ddv;
          if (l == j) ddv = vd[j];
          else { ddv = -v * qR[l]; v *= tR[l]; }
          double val = ph[l] * ph[j] * ddv;
          if (l == j) val += ie[l];
          emit(val); KMPC_FLUSH_IF_FULL();
        }
      }
      if (en % KB != 0) flush_batch(en - en % KB, en % KB);
    } else {
      if (en % KB != 0) flush_batch(en - en % KB, en % KB);
    }
#undef KMPC_FLUSH_IF_FULL
    }
  }

  // ---- border matrix on the fp64 tensor cores (see KDst): my asset's X | Y columns, X' Y over my warp's 32 assets,
  // warp 0 adds up the warps and scatters the needed products into the lower triangle of K -------------------------
  __device__ __forceinline__ void assemble_K_dmma(const double (&gjj)[H], const double (&vd)[H], const double (&fL)[H],
                                                  const double (&qL)[H], const double (&qR)[H], const double (&tR)[H]) {
    constexpr int HH = (H <= 5 ? H : 1);
    double* col = sm + OFF_TILE + tid;                       // column `tid` of the X | Y rows
    {
      double T = 1.0, Tprev = 1.0, iT = 1.0;
#pragma unroll
      for (int k = 0; k < H; ++k) {
        // no u variables (lam = 0 and tau <= 0): the stages are decoupled, tR = 0 exactly and T_l / T_j has no meaning;
        // a decay of 2^-200 per stage puts the off-diagonal blocks 60 decades below the diagonal ones instead of at 0
        if (k > 0) { Tprev = T; T *= hu() ? tR[k] : 0x1p-200; iT = rcp_fast(T); }
        const double gq = gjj[k] * iT;                       // padding lanes: gjj = vd = ie = 0 -> zero Y columns
        const double phk = hc() ? ph[k] : 0.0;
        col[k * LDX] = R[k] * T;                                            // P_R
        col[(H + k) * LDX] = T;                                             // P_1  (k = 0: the ones column)
        col[(2 * H + k) * LDX] = (k > 0) ? phk * qR[k] * Tprev : 0.0;       // Pe2
        double* y = col + 8 * MT * LDX;
        y[k * LDX] = R[k] * gq;                                             // Q_R
        y[(H + k) * LDX] = gq;                                              // Q_1
        y[(2 * H + k) * LDX] = -phk * qL[k] * gq;                           // Qe1
        if (k < H - 1) y[(3 * H + k) * LDX] = -phk * vd[k] * fL[k] * iT;    // Qe3
        // diagonal of the cap block; without a cap the rows stay decoupled with a positive diagonal (= N)
        y[(4 * H - 1 + k) * LDX] = hc() ? fma(phk * phk, vd[k], ie[k]) : (valid ? 1.0 : 0.0);
      }
    }
    __syncwarp();                                            // a warp contracts over its OWN 32 columns only
    double acc[MT][NTL][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) { acc[mt][nt][0] = 0.0; acc[mt][nt][1] = 0.0; }
    const double* frag = sm + OFF_TILE + (lane >> 2) * LDX + 32 * warp + (lane & 3);
#pragma unroll
    for (int st = 0; st < 8; ++st) {                         // 4 assets per step
      double av[MT], bv[NTL];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) av[mt] = frag[(8 * mt) * LDX + 4 * st];
#pragma unroll
      for (int nt = 0; nt < NTL; ++nt) bv[nt] = frag[(8 * (MT + nt)) * LDX + 4 * st];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NTL; ++nt) dmma_m8n8k4(acc[mt][nt][0], acc[mt][nt][1], av[mt], bv[nt]);
    }
    if (G > 1) {
      double* part = sm + OFF_TILE + XY_DOUBLES;
      if (warp != bw) {
        const int pw = warp - (warp > bw ? 1 : 0);           // my slot among the G - 1 partial products
#pragma unroll
        for (int q = 0; q < MT * NTL * 2; ++q) part[(pw * MT * NTL * 2 + q) * 32 + lane] = acc[q / (2 * NTL)][(q / 2) % NTL][q & 1];
      }
      sync();
    }
    if (warp == bw) {
      const double* part = sm + OFF_TILE + XY_DOUBLES;
#pragma unroll
      for (int q = 0; q < MT * NTL * 2; ++q) {
        double v = acc[q / (2 * NTL)][(q / 2) % NTL][q & 1];
#pragma unroll
        for (int g = 1; g < G; ++g) v += part[((g - 1) * MT * NTL * 2 + q) * 32 + lane];
        const int dst = g_kdst<HH>.dst[q][lane];
        if (dst >= 0) sm[OFF_K + dst] = v;
      }
    }
  }

  // L D L' factorisation of K by warp 0: lane = row, row in registers, columns exchanged by shuffles.  False on a
  // non-positive pivot.
  __device__ __forceinline__ bool factor_b() {
    if (warp == bw) {
      __syncwarp();                  // K was scattered by lanes of this warp (flush_batch); nobody else touches it
      const int nb = hc() ? 3 * H : 2 * H;
      const int r = (lane < NB) ? lane : NB - 1;
      if (lane < NB) {             // diagonal terms 1/beta_k = rho_k^2 and sc_k / zc_k, added in place (a lane-indexed
        double add = 0.0;          // update of the register row would push the whole row to local memory)
        if (lane < H) { const double rho = U(U_RHO, lane); add = rho * rho; }
        if (hc() && lane >= 2 * H) add = U(U_SC, lane - 2 * H) * rcp_fast(U(U_ZC, lane - 2 * H));
        sm[OFF_K + lane * NB + lane] += add;
      }
      __syncwarp();
      double a[NB];
#pragma unroll
      for (int c = 0; c < NB; ++c) a[c] = (c <= lane) ? sm[OFF_K + r * NB + c] : 0.0;
      // K = L D L' (unit lower L): no square root, and the column broadcast shfl(a[j], c) does not wait for 1/D_j
      bool pd = true;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const double aj = a[j];                              // row `lane`, column j, before scaling
        const double djj = shfl_d(aj, j);
        if (!(djj > 0.0) && j < nb) pd = false;
        const double inv = rcp_fast(djj);
        const double l = aj * inv;                           // L[lane][j] (lanes > j)
        a[j] = l;
        if (lane == j) sm[OFF_K + NB * NB + j] = inv;
#pragma unroll
        for (int c = j + 1; c < NB; ++c) a[c] = fma(-l, shfl_d(aj, c), a[c]);
      }
      if (lane < NB) {
#pragma unroll
        for (int c = 0; c < NB; ++c) if (c <= lane) sm[OFF_K + lane * NB + c] = a[c];
      }
      if (lane == 0) sm[OFF_FLAG] = pd ? 1.0 : 0.0;
    }
    sync();
    fact_ok_ = uni(sm[OFF_FLAG] > 0.5);
    return fact_ok_;
  }

  // OFF_T <- K^{-1} t by the border warp; lane r < NB passes entry r of the right-hand side in t_in.
  // The substitutions run in blocks of H unknowns: the block's right-hand sides are broadcast with H independent
  // shuffles, every lane solves the H x H unit-triangular diagonal block for itself (its entries are warp-uniform
  // loads) and then updates its own entry with the H solved values.  Same arithmetic as the plain substitution — one
  // shuffle round trip per BLOCK on the critical path instead of one per unknown (42 cycles each: 30 of them per solve).
#ifndef KMPC_KSOLVE_BLOCKED
#define KMPC_KSOLVE_BLOCKED 1
#endif
  __device__ __forceinline__ void k_solve_shared(double t_in) {
    if (warp == bw) {
      const int r = (lane < NB) ? lane : NB - 1;
      double Lr[NB], Lc[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        Lr[j] = (j < lane && lane < NB) ? sm[OFF_K + r * NB + j] : 0.0;          // L[lane][j]
        Lc[j] = (j > lane && lane < NB) ? sm[OFF_K + j * NB + r] : 0.0;          // L[j][lane]
      }
      const double myinv = (lane < NB) ? sm[OFF_K + NB * NB + r] : 0.0;
      double t = (lane < NB) ? t_in : 0.0;
#if KMPC_KSOLVE_BLOCKED
      constexpr int BS = H;
#pragma unroll
      for (int b0 = 0; b0 < NB; b0 += BS) {                                      // L y = t
        double y[BS];
#pragma unroll
        for (int i = 0; i < BS; ++i) y[i] = shfl_d(t, b0 + i);
#pragma unroll
        for (int i = 1; i < BS; ++i)
#pragma unroll
          for (int k = 0; k < i; ++k) y[i] = fma(-sm[OFF_K + (b0 + i) * NB + b0 + k], y[k], y[i]);
#pragma unroll
        for (int i = 0; i < BS; ++i) t = fma(-Lr[b0 + i], y[i], t);              // Lr[j] = 0 for j >= lane
      }
      t *= myinv;                                                                // D z = y
#pragma unroll
      for (int b0 = NB - BS; b0 >= 0; b0 -= BS) {                                // L' x = z
        double x[BS];
#pragma unroll
        for (int i = 0; i < BS; ++i) x[i] = shfl_d(t, b0 + i);
#pragma unroll
        for (int i = BS - 2; i >= 0; --i)
#pragma unroll
          for (int k = BS - 1; k > i; --k) x[i] = fma(-sm[OFF_K + (b0 + k) * NB + b0 + i], x[k], x[i]);
#pragma unroll
        for (int i = BS - 1; i >= 0; --i) t = fma(-Lc[b0 + i], x[i], t);         // Lc[j] = 0 for j <= lane
      }
#else
#pragma unroll
      for (int j = 0; j < NB; ++j) t = fma(-Lr[j], shfl_d(t, j), t);        // L y = t   (Lr[j] = 0 for j >= lane)
      t *= myinv;                                                              // D z = y
#pragma unroll
      for (int j = NB - 1; j >= 0; --j) t = fma(-Lc[j], shfl_d(t, j), t);   // L' x = z  (Lc[j] = 0 for j <= lane)
#endif
      if (lane < NB) sm[OFF_T + lane] = t;
    }
  }

  // One Newton solve; on return dw/dsp/dsq/dzw/dzp/dzq hold the direction of my asset, the stage scalars
  // (dnu, dsc, dzc) of stage tid are returned to the owner threads tid < H, and (rp_, rd_) are my largest primal / dual step ratios -dv/v.
  __device__ __forceinline__ void newton(bool use_c, double (&dw)[H], double (&dsp)[H], double (&dsq)[H],
                                         double (&dzw)[H], double (&dzp)[H], double (&dzq)[H], double& dnu, double& dsc,
                                         double& dzc, int& rp_, int& rd_) {
    double gw[H], gu[H], pg[H], tq[H];
#pragma unroll
    for (int k = 0; k < H; ++k) {
      double gwk = fma(R[k], U(U_IRHO, k), -U(U_NU, k));
      double guk = 0.0;
      tq[k] = 0.0;
      if (use_c && hw()) gwk = fma(TGT(T_CW, k), iw[k], gwk);
      if (hu()) {
        double a1 = 0.0, a2 = 0.0;
        if (use_c) { a1 = TGT(T_CP, k) * isp[k]; a2 = TGT(T_CQ, k) * isq[k]; }
        tq[k] = a1 - a2;
        guk = -lam + a1 + a2;
        if (hc()) guk = fma(-U(U_CC, k), U(U_ISC, k), guk);
      }
      gw[k] = gwk; gu[k] = guk;
    }
#pragma unroll
    for (int k = 0; k < H; ++k) {
      gw[k] -= tq[k];
      if (k + 1 < H) gw[k] += tq[(k + 1 < H) ? k + 1 : 0];
      pg[k] = ph[k] * gu[k];
    }
    double dd[H];
    m0_apply(gw, pg, dw, dd);
    {
      double v[NB];
#pragma unroll
      for (int k = 0; k < H; ++k) {
        v[k] = R[k] * dw[k];                                 // padding lanes: dw = dd = ie = 0
        v[H + k] = dw[k];
        v[2 * H + k] = hc() ? fma(gu[k], ie[k], -ph[k] * dd[k]) : 0.0;
      }
      tile_reduce<NB>(v);
      double t = 0.0;
      if (warp == bw && lane < NB) {
        t = ptotal(lane);
        if (lane >= H && lane < 2 * H) t += U(U_RP, lane - H);  // t[H+k] = sum dw0 - q, q = -rp
      }
      KMPC_PROF(*this, 7)
      k_solve_shared(t);
    }
    sync();
    KMPC_PROF(*this, 8)
    dnu = 0.0; dsc = 0.0; dzc = 0.0;
    if (tid < H) {                                           // stage scalars stay with their owner threads
      const double yC = hc() ? sm[OFF_T + 2 * H + tid] : 0.0;
      dnu = sm[OFF_T + H + tid];
      if (hc()) {
        dsc = -yC * U(U_SC, tid) * rcp_fast(U(U_ZC, tid));
        dzc = fma(U(U_CC, tid), U(U_ISC, tid), -U(U_ZC, tid)) + yC;
      }
    }
    double geff[H];
#pragma unroll
    for (int k = 0; k < H; ++k) {
      gw[k] -= fma(sm[OFF_T + k], R[k], sm[OFF_T + H + k]);
      geff[k] = gu[k] - (hc() ? sm[OFF_T + 2 * H + k] : 0.0);
      pg[k] = ph[k] * geff[k];
    }
    m0_apply(gw, pg, dw, dd);
    int rp = 0, rd = 0;                                     // high words of the largest ratios -dv/v
    if (valid) {
#pragma unroll
      for (int k = 0; k < H; ++k) {
        if (hw()) {
          const double cw = use_c ? TGT(T_CW, k) : 0.0;
          dzw[k] = fma(iw[k], fma(-zw[k], dw[k], cw), -zw[k]);
          rp = max(rp, hi_of(-dw[k] * iw[k]));
          rd = max(rd, hi_of(-dzw[k] * rcp_fast(zw[k])));
        } else dzw[k] = 0.0;
        if (hu()) {
          const double cp = use_c ? TGT(T_CP, k) : 0.0, cq = use_c ? TGT(T_CQ, k) : 0.0;
          const double dp = zp[k] * isp[k], dq = zq[k] * isq[k];
          dsp[k] = fma(-fma(2.0, dq, delta), dd[k], geff[k]) * ie[k];
          dsq[k] = fma(fma(2.0, dp, delta), dd[k], geff[k]) * ie[k];
          dzp[k] = fma(isp[k], fma(-zp[k], dsp[k], cp), -zp[k]);
          dzq[k] = fma(isq[k], fma(-zq[k], dsq[k], cq), -zq[k]);
          rp = max(rp, max(hi_of(-dsp[k] * isp[k]), hi_of(-dsq[k] * isq[k])));
          rd = max(rd, max(hi_of(-dzp[k] * rcp_fast(zp[k])), hi_of(-dzq[k] * rcp_fast(zq[k]))));
        } else { dsp[k] = 0.0; dsq[k] = 0.0; dzp[k] = 0.0; dzq[k] = 0.0; }
      }
    } else {
#pragma unroll
      for (int k = 0; k < H; ++k) { dw[k] = 0.0; dsp[k] = 0.0; dsq[k] = 0.0; dzw[k] = 0.0; dzp[k] = 0.0; dzq[k] = 0.0; }
    }
    if (hc() && tid < H) {
      rp = max(rp, hi_of(-dsc * U(U_ISC, tid)));
      rd = max(rd, hi_of(-dzc * rcp_fast(U(U_ZC, tid))));
    }
    rp_ = rp; rd_ = rd;
  }

  // R[k] = exp(y[k * stride + my asset]) in float32 like the reference (mpc.py:55), plus exp(y_extra) (the caller's
  // realised return, backtest.py:193).  All loads are issued first; ONE rolled copy of the exp code then works through
  // my column of the reduction tile (free between reductions; nobody else touches my column): see div_fast for why
  // the code is kept small.  Call with the tile idle, i.e. after the previous reduction's results have been consumed.
  // `ycol` = my asset's column of y (tid, except for the compacted problems of the active-set kernel).
  __device__ __forceinline__ float load_returns(const float* y, size_t stride, float y_extra, int ycol) {
    static_assert(H + 1 <= SMALL_ROWS, "tile too small");
    double* col = sm + OFF_TILE + tid;
    float x[H];
#pragma unroll
    for (int k = 0; k < H; ++k) x[k] = valid ? y[(size_t)k * stride + ycol] : 0.0f;
#pragma unroll
    for (int k = 0; k < H; ++k) col[k * LD] = (double)x[k];
    col[H * LD] = (double)y_extra;
#pragma unroll 1
    for (int k = 0; k <= H; ++k) col[k * LD] = (double)__double2float_rn(exp(col[k * LD]));
#pragma unroll
    for (int k = 0; k < H; ++k) R[k] = col[k * LD];
    return (float)col[H * LD];
  }

  // ---- the solve, in pieces (so that a caller can interleave several problems in lockstep) ---------------------
  //   begin()        screening + initial point.  Returns -1 (iterate) or a terminal status.
  //   check()        residuals of the current iterate.  Returns -1 (take another Newton step) or the final status.
  //   factor_a/b()   factorisation at the current iterate
  //   newton_phase() predictor (0) and corrector + step (1)
  // R[] (gross returns of my asset, all stages) must be set before begin(); w0 = my current weight.
  // restart = true: the second attempt of the same problem (check() returned ST_RESTART): same R, same w0, the
  // iteration count carries on.
  __device__ __forceinline__ int begin(double w0, int N, double lam_, double tau_, bool allow_short,
                                       const IpmOptions& opt, bool restart = false) {
    robust_ = uni(restart);
    lam = lam_; tau = tau_; delta = robust_ ? kRobustDelta : opt.delta;
    has_u_ = uni((lam > 0.0) || (tau > 0.0));
    has_c_ = has_u_ && uni(tau > 0.0);
    has_w_ = !allow_short; allow_short_ = allow_short;
    if (!robust_) it_ = 0;
    it0_ = it_; fact_ok_ = true; nretry_ = 0;
    kkt_[0] = kkt_[1] = kkt_[2] = CUDART_NAN;
    if (!valid) {
      w0 = 0.0;
#pragma unroll
      for (int k = 0; k < H; ++k) R[k] = 1.0;
    }
    w0_ = w0;
    // ---- screening, initial point (oracle/mpc_oracle.py::_initial_point) --------------------------------------
    const double base = valid ? (ash() ? w0 : fmax(w0, 0.0)) : 0.0;
    double mxR[H];
    double sb;
    {   // one round: stage maxima of R, sum of the clipped weights, count of non-finite inputs
      bool okv = isfinite(w0);
#pragma unroll
      for (int k = 0; k < H; ++k) okv = okv && isfinite(R[k]) && (R[k] > 0.0);
      double v[H + 2];
#pragma unroll
      for (int k = 0; k < H; ++k) v[k] = (valid && okv) ? R[k] : 0.0;
      v[H] = (valid && okv) ? base : 0.0;
      v[H + 1] = okv ? 0.0 : 1.0;
      tile_reduce_mixed<H, H + 2>(v);
      if (uni(ptotal(H + 1) > 0.0)) return finish(ST_NONFINITE);
#pragma unroll
      for (int k = 0; k < H; ++k) mxR[k] = ptotal_max(k);
      sb = ptotal(H);
    }
    const double invN = rcp_fast((double)N);
    const double eps = (tau <= 0.0) ? 0.1 : fmin(0.1, tau * 0.125);
    const double b0 = (sb > 0.0) ? div_fast(base, sb) : invN;
    const double w1 = valid ? fma(1.0 - eps, b0, eps * invN) : 1.0;
    double rho0[H];
    double absd0;
    {
      double v[H + 1], tot[H + 1];
#pragma unroll
      for (int k = 0; k < H; ++k) v[k] = valid ? w1 * R[k] : 0.0;
      v[H] = valid ? fabs(w1 - w0) : 0.0;
      block_sum<H + 1>(v, tot);
#pragma unroll
      for (int k = 0; k < H; ++k) rho0[k] = tot[k];
      absd0 = tot[H];
    }
    double sc0 = 1.0, sck = 1.0, dl0 = 0.0, dlk = 0.0;
    if (hu()) {
      if (uni(tau > 0.0)) {
        const double room0 = tau - absd0;
        if (uni(!(room0 > 0.0))) {
          kkt_[0] = kkt_[1] = kkt_[2] = CUDART_INF;
          return finish(ST_FAILED);
        }
        dl0 = 0.5 * room0 * invN; dlk = 0.5 * tau * invN;
      } else { dl0 = dlk = 0.05 * invN; }
      if (hc()) { sc0 = tau - (absd0 + dl0 * N); sck = tau - dlk * N; }
    }
    // FIX kernels are only launched with dual_init > 0 (lane_fix_plan): the mu0-based start is compiled out there
    const bool dual_start = FIX ? true : (hw() && uni(opt.dual_init > 0.0));
    const double zeta0 = hc() ? opt.dual_init : 0.0;
#pragma unroll
    for (int k = 0; k < H; ++k) {
      const double d0 = (k == 0 && valid) ? w1 - w0 : 0.0;
      const double uk = hu() ? ((k == 0) ? fabs(d0) + dl0 : dlk) : 1.0;
      w[k] = w1;
      sp[k] = hu() ? uk - d0 : 1.0;
      sq[k] = hu() ? uk + d0 : 1.0;
      const double sck_ = (k == 0) ? sc0 : sck;
      double nu_k;
      if (dual_start) {
        const double ir = rcp_fast(rho0[k]);
        nu_k = mxR[k] * ir + opt.dual_init;
        zw[k] = valid ? fma(-R[k], ir, nu_k) : 0.0;
        zp[k] = hu() ? 0.5 * fmax(lam + zeta0, opt.dual_init) : 0.0;   // floor for the uncapped case: see oracle
        zq[k] = zp[k];
        if (tid == 0) U(U_ZC, k) = hc() ? zeta0 : 0.0;
      } else {
        nu_k = 1.0;
        zw[k] = (hw() && valid) ? div_fast(opt.mu0, w[k]) : 0.0;
        zp[k] = hu() ? div_fast(opt.mu0, sp[k]) : 0.0;
        zq[k] = hu() ? div_fast(opt.mu0, sq[k]) : 0.0;
        if (tid == 0) U(U_ZC, k) = hc() ? div_fast(opt.mu0, sck_) : 0.0;
      }
      if (tid == 0) { U(U_NU, k) = nu_k; U(U_SC, k) = sck_; U(U_CC, k) = 0.0; }
      if (!valid) { zw[k] = 1.0; zp[k] = 1.0; zq[k] = 1.0; }      // benign padding (never updated, never summed)
    }
    sync();
    mcount_ = (hw() ? (double)H * N : 0.0) + (hu() ? 2.0 * H * N : 0.0) + (hc() ? (double)H : 0.0);
    return -1;
  }

  // An attempt ends without "optimal": the first attempt hands over to the second one (the caller runs begin() again
  // with restart = true), the second one ends the solve.
  __device__ __forceinline__ int end_attempt(int status, const IpmOptions& opt) {
    if (!robust_ && opt.second_attempt && status != ST_NONFINITE) return ST_RESTART;
    return finish(status);
  }
  // final status of a solve that left the iteration without meeting the tolerances (or never started)
  __device__ __forceinline__ int finish(int status) {
    if (status == ST_FAILED && isfinite(kkt_[1] + kkt_[2]) && kkt_[0] < kLoosePres && kkt_[1] < kLooseDres && kkt_[2] < kLooseGap)
      status = ST_INACCURATE;
    if (status >= ST_FAILED) {
#pragma unroll
      for (int k = 0; k < H; ++k) w[k] = w0_;              // mpc.py:113-115: hold the current weights
    }
    return status;
  }

  __device__ __forceinline__ int check(const IpmOptions& opt) {
    if (!fact_ok_) {
      // The border factorisation met a non-positive pivot (barrier weights spanning > 20 decades next to the optimum,
      // usually after the endgame has shrunk delta).  The iterate is untouched: retry a few times with a stronger
      // proximal term — a damped but well-conditioned Newton step — before giving up.
      // The first breakdown is always retried (most such iterates then reach the tolerances); from the second one on
      // an iterate that meets the loose bar is accepted as it is (finish() applies the loose acceptance).
      const bool loose = isfinite(kkt_[1] + kkt_[2]) && kkt_[0] < kLoosePres && kkt_[1] < kLooseDres && kkt_[2] < kLooseGap;
      if (nretry_ >= kMaxFactorRetries || (nretry_ > 0 && loose)) return end_attempt(ST_FAILED, opt);
      ++nretry_;
      delta = fmin(fmax(delta, robust_ ? kRobustDelta : opt.delta) * 30.0, 1e-2);
      fact_ok_ = true;
    }
    ++it_;
    double gap, pres;
    {
      double v[2 * H + 1];
      double g = 0.0;
#pragma unroll
      for (int k = 0; k < H; ++k) {
        v[k] = valid ? w[k] * R[k] : 0.0;
        v[H + k] = valid ? w[k] : 0.0;
        if (hw()) g = fma(w[k], zw[k], g);
        if (hu()) g = fma(sp[k], zp[k], fma(sq[k], zq[k], g));
      }
      if (!valid) g = 0.0;
      if (hc() && tid < H) g = fma(U(U_SC, tid), U(U_ZC, tid), g);
      v[2 * H] = g;
      tile_reduce<2 * H + 1>(v);
      gap = ptotal(2 * H);
      pres = 0.0;
#pragma unroll
      for (int k = 0; k < H; ++k) pres = fmax(pres, fabs(ptotal(H + k) - 1.0));
      if (tid < H) {
        const double rho = ptotal(tid);
        U(U_RHO, tid) = rho; U(U_IRHO, tid) = rcp_fast(rho);
        U(U_RP, tid) = ptotal(H + tid) - 1.0;
        U(U_ISC, tid) = hc() ? rcp_fast(U(U_SC, tid)) : 0.0;
      }
    }
    sync();
    // The dual residual only decides anything once the gap is small (every acceptance test below, and the loose one
    // in finish(), asks for gap < max(tol, kLooseGap)) or at the iteration cap: it is not evaluated before that (NaN).
    const bool last_it = (it_ - it0_ == opt.max_iter + 1);
    const bool near = uni(gap < fmax(opt.tol, kLooseGap)) || last_it;
    double dres = CUDART_NAN;
    if (near) {
      int dres_h = 0;
      if (valid) {
#pragma unroll
        for (int k = 0; k < H; ++k) {
          const double yk = zp[k] - zq[k];
          const double yn = (k + 1 < H) ? zp[(k + 1 < H) ? k + 1 : 0] - zq[(k + 1 < H) ? k + 1 : 0] : 0.0;
          const double rdw = fma(-R[k], U(U_IRHO, k), U(U_NU, k)) - (hw() ? zw[k] : 0.0) + (yk - yn);
          dres_h = max(dres_h, hi_of(fabs(rdw)));
          if (hu()) dres_h = max(dres_h, hi_of(fabs(lam - zp[k] - zq[k] + (hc() ? U(U_ZC, k) : 0.0))));
        }
      }
      int dummy = 0;
      block_max2i(dres_h, dummy);
      dres = (dres_h > 0) ? hi_up(dres_h) : 0.0;      // a NaN residual has a large positive high word: +inf-like
    }
    kkt_[0] = pres; kkt_[1] = dres; kkt_[2] = gap;
    gap_ = gap;
    if (uni(!isfinite(gap) || (near && !isfinite(dres)))) return end_attempt(ST_FAILED, opt);
    if (uni(pres < opt.tol && dres < opt.tol_dual && gap < opt.tol)) return ST_OPTIMAL;
    // flat directions (curvature << delta): the dual residual crawls at ~delta*|dx| while the gap has long
    // collapsed; the objective is converged -> "optimal_inaccurate" instead of iterating into round-off
    if (uni(pres < opt.tol && gap < 1e-6 * opt.tol && dres < 1e-6)) {
      if (!robust_ && opt.second_attempt) return ST_RESTART;
      return ST_INACCURATE;
    }
    if (last_it) return end_attempt(ST_FAILED, opt);
    mu_ = div_fast(gap, fmax(mcount_, 1.0));
    if (uni(pres < opt.tol && gap < opt.tol)) delta = fmax(0.3 * delta, 1e-9);   // endgame: shrink the proximal term
    return -1;
  }

  // phase 0: predictor -> complementarity targets of the corrector; phase 1: corrector -> step
  __device__ __forceinline__ void newton_phase(int phase, const IpmOptions& opt) {
    const bool has_m = FIX ? true : (mcount_ > 0.0);
    if (phase == 0 && !has_m) return;
    const bool use_c = (phase == 1) && has_m;
    const bool stepped = has_m || ash();
    double dw[H], dsp[H], dsq[H], dzw[H], dzp[H], dzq[H], dnu, dsc, dzc;
    int rph, rdh;
    newton(use_c, dw, dsp, dsq, dzw, dzp, dzq, dnu, dsc, dzc, rph, rdh);
    if (ash()) {            // keep the argument of the logarithm positive: rho_k + a * sum_i R dw > 0
      double v[H], tot[H];
#pragma unroll
      for (int k = 0; k < H; ++k) v[k] = dw[k] * R[k];
      block_sum<H>(v, tot);
#pragma unroll
      for (int k = 0; k < H; ++k) rph = max(rph, hi_of(-tot[k] * U(U_IRHO, k)));
    }
    block_max2i(rph, rdh);
    const double rp = (rph > 0) ? hi_up(rph) : 0.0, rd = (rdh > 0) ? hi_up(rdh) : 0.0;
    // largest steps keeping slacks (aa) and duals (ab) non-negative: min(1, 1 / max ratio)
    const double aa = (stepped && rp > 1.0) ? rcp_fast(rp) : 1.0;
    const double ab = (stepped && rd > 1.0) ? rcp_fast(rd) : 1.0;
    if (phase == 0) {
      double g2 = 0.0;
      if (valid) {
#pragma unroll
        for (int k = 0; k < H; ++k) {
          if (hw()) g2 = fma(fma(aa, dw[k], w[k]), fma(ab, dzw[k], zw[k]), g2);
          if (hu()) g2 = fma(fma(aa, dsp[k], sp[k]), fma(ab, dzp[k], zp[k]),
                              fma(fma(aa, dsq[k], sq[k]), fma(ab, dzq[k], zq[k]), g2));
        }
      }
      if (hc() && tid < H) g2 = fma(fma(aa, dsc, U(U_SC, tid)), fma(ab, dzc, U(U_ZC, tid)), g2);
      g2 = block_sum1(g2);
      const double ratio = (gap_ > 0.0) ? fmin(1.0, fmax(div_fast(g2, gap_), 0.0)) : 0.0;
      const double smu = fmax(ratio * ratio * ratio, robust_ ? kRobustSigmaMin : 0.0) * mu_;
      const double dmp = fmin(1.0, fmin(aa, ab) * (1.0 / kCorrFull));   // short affine step: damp the corrector
#pragma unroll
      for (int k = 0; k < H; ++k) {      // complementarity targets of the corrector
        TGT(T_CW, k) = hw() ? fma(-dmp * dw[k], dzw[k], smu) : 0.0;
        TGT(T_CP, k) = hu() ? fma(-dmp * dsp[k], dzp[k], smu) : 0.0;
        TGT(T_CQ, k) = hu() ? fma(-dmp * dsq[k], dzq[k], smu) : 0.0;
      }
      if (tid < H) U(U_CC, tid) = hc() ? fma(-dmp * dsc, dzc, smu) : 0.0;
      sync();
    } else {
      double pa = stepped ? fmin(1.0, opt.step_frac * aa) : 1.0;
      double pb = stepped ? fmin(1.0, opt.step_frac * ab) : 1.0;
      if (robust_) { pa = stepped ? fmin(1.0, kRobustStepFrac * fmin(aa, ab)) : 1.0; pb = pa; }
      if (valid) {
#pragma unroll
        for (int k = 0; k < H; ++k) {
          w[k] = fma(pa, dw[k], w[k]);
          if (hw()) zw[k] = fma(pb, dzw[k], zw[k]);
          if (hu()) {
            sp[k] = fma(pa, dsp[k], sp[k]); sq[k] = fma(pa, dsq[k], sq[k]);
            zp[k] = fma(pb, dzp[k], zp[k]); zq[k] = fma(pb, dzq[k], zq[k]);
          }
        }
      }
      if (tid < H) {
        U(U_NU, tid) = fma(pb, dnu, U(U_NU, tid));
        if (hc()) { U(U_SC, tid) = fma(pa, dsc, U(U_SC, tid)); U(U_ZC, tid) = fma(pb, dzc, U(U_ZC, tid)); }
        U(U_CC, tid) = 0.0;                  // the next predictor has no complementarity targets
      }
      sync();
    }
  }

  // Solve one problem start to end.  Returns the status; w[] holds my entries of the plan (w0 in every stage on
  // failure), kkt = (primal residual, dual residual, complementarity gap) of the last iterate.
  __device__ __forceinline__ int solve(double w0, int N, double lam_, double tau_, bool allow_short,
                                       const IpmOptions& opt, int& iters, double (&kkt)[3]) {
    int status = begin(w0, N, lam_, tau_, allow_short, opt);
    while (status < 0) {
      if (status == ST_RESTART) {                      // second attempt: same R, same w0
        sync();
        status = begin(w0, N, lam_, tau_, allow_short, opt, true);
        continue;
      }
      status = check(opt);
      if (status >= 0 || status == ST_RESTART) continue;
      factor_a();
      if (!factor_b()) continue;
#pragma unroll 1
      for (int phase = 0; phase < 2; ++phase) newton_phase(phase, opt);
    }
    iters = it_;
    kkt[0] = kkt_[0]; kkt[1] = kkt_[1]; kkt[2] = kkt_[2];
    return status;
  }

  // The first trade of the plan is the one the caller executes (backtest.py:131).  An iterate accepted after the
  // factorisation broke down next to the optimum can sit ~1e-5 outside the turnover cap (the cap's slack is stepped,
  // not re-derived from w): pull that trade back onto the cap along its own direction.
  __device__ __forceinline__ void clip_first_trade(double w0) {
    if (!uni(tau > 0.0)) return;
    sync();
    const double t = block_sum1(valid ? fabs(w[0] - w0) : 0.0);
    if (t > tau && valid) w[0] = fma(div_fast(tau, t), w[0] - w0, w0);
  }

  // maximised objective (mpc.py:104) of the plan held in w[]; same value in every thread
  __device__ __forceinline__ double objective(double w0) {
    double v[H + 1], tot[H + 1];
    double ab = 0.0;
#pragma unroll
    for (int k = 0; k < H; ++k) {
      v[k] = valid ? w[k] * R[k] : 0.0;
      ab += valid ? fabs(w[k] - ((k == 0) ? w0 : w[(k == 0) ? 0 : k - 1])) : 0.0;
    }
    v[H] = ab;
    sync();
    block_sum<H + 1>(v, tot);
    double val = -lam * tot[H];
#pragma unroll
    for (int k = 0; k < H; ++k) val += log(tot[k]);
    return val;
  }
};

}  // namespace kmpc
