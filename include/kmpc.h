/* libkmpc — C ABI of the B200-native Koopman-forecast + MPC rebalancing hot path.
 *
 * Drop-in boundary for the data-parallel hot path of yli421/koopman-mpc-portfolio-rebalancing.  The
 * reference is pure Python and has no FFI of its own; its seams are Python call signatures
 * (SURVEY.md §8b).  Each entry point below names the reference code it replaces; INTEGRATION.md shows
 * the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C, no exceptions, no allocation handed across the ABI except opaque handles;
 *   - every array argument is a DEVICE pointer unless its name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous
 *     on that stream unless stated otherwise;
 *   - return value: 0 = success, negative = KMPC_E_*; kmpc_last_error() gives a message
 *     (thread-local);
 *   - a handle is bound to one device; not thread-safe per handle, re-entrant across handles.  A handle owns a work
 *     counter, device flags and scratch buffers: issue its calls on ONE stream at a time (calls on one stream are
 *     ordered and safe back to back); concurrent streams or threads need a handle each.  Entry points leave the
 *     caller's current CUDA device unchanged.
 *   - there is NO CPU fallback: every entry point fails with KMPC_E_CUDA / KMPC_E_UNSUPPORTED if the
 *     device or a compiled kernel variant is missing.
 */
#ifndef KMPC_H_
#define KMPC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMPC_VERSION 100

enum {
  KMPC_OK = 0,
  KMPC_E_INVALID = -1,     /* bad argument */
  KMPC_E_UNSUPPORTED = -2, /* shape without a compiled kernel variant (e.g. N > 128, H not built) */
  KMPC_E_CUDA = -3,        /* CUDA runtime error */
  KMPC_E_NOMEM = -4
};

/* per-problem solver status (mpc.py:113-117 status strings in parentheses) */
enum {
  KMPC_STATUS_OPTIMAL = 0,    /* "optimal" */
  KMPC_STATUS_INACCURATE = 1, /* "optimal_inaccurate": loose tolerances met, weights returned */
  KMPC_STATUS_FAILED = 2,     /* "solver_error": fallback tile(w_cur), value None (mpc.py:113-115) */
  KMPC_STATUS_NONFINITE = 3   /* non-finite / non-positive input: same fallback */
};

/* activation of MLPCoder layers, model.py:43-59 */
enum { KMPC_ACT_RELU = 0, KMPC_ACT_TANH = 1, KMPC_ACT_GELU = 2 };
/* model families, model.py:878-882 */
enum { KMPC_MODEL_GENERIC = 0 /* GenericKM / SparseKM */, KMPC_MODEL_LISTA = 1 /* LISTAKM */ };
enum { KMPC_NORM_ID = 0, KMPC_NORM_BALL = 1 }; /* model.py:749-752 */

typedef struct kmpc_handle kmpc_handle;
typedef struct kmpc_model kmpc_model;

int kmpc_version(void);
const char* kmpc_last_error(void);

/* One handle per (device, caller).  Owns a small device workspace and the SM count. */
int kmpc_create(int device, kmpc_handle** out);
int kmpc_destroy(kmpc_handle* h);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t kmpc_launch_count(const kmpc_handle* h);
/* 1 if an MPC kernel variant for (H, N) is compiled in */
int kmpc_mpc_supported(int H, int N);
/* Solver options of this handle (defaults: csrc/mpc_common.cuh default_ipm_options).  Tuning / diagnostics only:
 * the parity tests run with the defaults.  KMPC_PARAM_DUAL_INIT = 0 selects the mu0-based starting point. */
enum {
  KMPC_PARAM_RESET = 0,
  KMPC_PARAM_STEP_FRAC = 1,
  KMPC_PARAM_DUAL_INIT = 2,
  KMPC_PARAM_MAX_ITER = 3,
  /* 1 [default]: an optimal_inaccurate plan whose first trade sits outside the turnover cap (by <= 2e-5) is scaled back
   * onto it */
  KMPC_PARAM_CLIP_FIRST_TRADE = 4,
  /* 1 [default]: a solve whose first attempt does not end "optimal" is repeated from the cold start with robust
   * parameters (csrc/mpc_common.cuh kRobust*) */
  KMPC_PARAM_SECOND_ATTEMPT = 5,
  /* tolerance on the complementarity gap and the primal residual (default: default_ipm_options().tol) */
  KMPC_PARAM_TOL = 6,
  /* 1 [default]: kmpc_backtest_run solves REDUCED problems once a backtest's portfolio has concentrated (32 < N <= 512,
   * H <= 5 or H = 10, long-only): only the held assets and the best forecasts of each stage enter the interior-point solve (one warp
   * per problem instead of two or four; eight instead of sixteen for universes beyond 128 assets), and the optimality conditions of every excluded asset are then checked
   * against the duals of the reduced solution (an asset that fails joins the set and the problem is solved again), so
   * the plan is an optimum of the FULL program of mpc.py:49-104.  0: every decision solves all N assets.  2 (test hook):
 * as 1 but the set starts from the held assets alone, so that the check-and-repair path does the selecting.  3 (tuning):
 * as 1 with whole backtests instead of 32-decision segments as the work items of the reduced-solve kernel; >= 8 (tuning):
 * as 1 with that many decisions per work item. */
  KMPC_PARAM_ACTIVE_SET = 7
};
int kmpc_set_solver_param(kmpc_handle* h, int which, double value);

/* ---------------------------------------------------------------------------------------------
 * Data side — replaces data_finance.py
 * ------------------------------------------------------------------------------------------- */

/* standardize_returns (data_finance.py:243-259) + the float32 cast of create_finance_splits (:331):
 * out[b,t,a] = (float)((logret[b,t,a] - mean[b',a]) / std[b',a]), fp64 arithmetic, bit-exact.
 * mean/std are [B,N] (per path) when stats_per_path != 0, else [N] shared.  out row stride ld_out >= N;
 * padding columns are written as 0. */
int kmpc_standardize(kmpc_handle* h, const double* logret, const double* mean, const double* std,
                     int stats_per_path, int B, int T, int N, float* out, int ld_out, void* stream);

/* time_delay_embedding (data_finance.py:262-300), materialised (the forecast kernels read the window
 * in place and never need this; it exists for FinanceDataset.data and for parity checks):
 * out[b,i,j*N+a] = data[b,i+d-1-j,a],  i in [0,T-d+1), data row stride ld_in. */
int kmpc_embed_gather(kmpc_handle* h, const float* data, int ld_in, int B, int T, int N, int d, float* out,
                      void* stream);

/* The gather map itself, host side (int32 [T-d+1, d*N], index into data.ravel() with ld = N).
 * Returns KMPC_E_INVALID when T < d (the reference raises ValueError, data_finance.py:281-282). */
int kmpc_embed_index_host(int T, int N, int d, int32_t* idx_out_host);

/* FinanceEnv.extract_current_returns + destandardize_returns (data_finance.py:717-742) for every test row:
 * out[b,r,a] = fl32(fl32(z[b,row0+r+d-1,a] * (float)std[a]) + (float)mean[a])   (two roundings, no FMA) */
int kmpc_current_returns(kmpc_handle* h, const float* z, int ld_z, const double* mean, const double* std,
                         int stats_per_path, int B, int T, int N, int d, int row0, int rows, float* out,
                         void* stream);

/* ---------------------------------------------------------------------------------------------
 * Forecast — replaces model.py encode/step_latent/decode + backtest.py:85-121
 * ------------------------------------------------------------------------------------------- */

typedef struct kmpc_model_desc {
  int kind;           /* KMPC_MODEL_* */
  int obs;            /* observation size d*N */
  int n_assets;       /* N */
  int delay;          /* d */
  int latent;         /* TARGET_SIZE Z */
  int norm_fn;        /* KMPC_NORM_* (GenericKM only) */
  /* encoder MLP (GenericKM.encoder or LISTA.We when lista_linear_encoder == 0): n_enc Linear layers */
  int n_enc;
  const int* enc_dims_host;          /* [n_enc+1] host: obs, h1, ..., Z */
  const float* const* enc_w_host;    /* host array of n_enc DEVICE pointers, weight [out,in] row-major */
  const float* const* enc_b_host;    /* host array of n_enc DEVICE pointers or NULL entries (no bias) */
  int enc_act;        /* KMPC_ACT_* */
  int enc_last_relu;
  /* decoder MLP (GenericKM): n_dec Linear layers Z -> ... -> obs */
  int n_dec;
  const int* dec_dims_host;
  const float* const* dec_w_host;
  const float* const* dec_b_host;
  int dec_act;
  const float* kmat;  /* [Z,Z] device, row-vector convention z_{k+1} = z_k @ kmat (model.py:320-321) */
  /* LISTAKM (model.py:120-209, 801-850) */
  int lista_linear_encoder;
  const float* lista_We;   /* [Z,obs] device when lista_linear_encoder */
  const float* lista_S;    /* [Z,Z] device, used as z @ S */
  const float* lista_dict; /* [Z,obs] device; decode = z @ (dict / ||dict||_row.clamp(1e-4)) */
  int lista_loops;
  float lista_threshold;   /* alpha / L */
} kmpc_model_desc;

/* Copies/re-lays the weights into the handle's own device buffers (column permutation for the in-place
 * window read, row-normalised dictionary, first-N decoder rows).  Synchronous. */
int kmpc_model_load(kmpc_handle* h, const kmpc_model_desc* desc, kmpc_model** out);
int kmpc_model_free(kmpc_model* m);

/* Forecasts for rows t0 <= t < t1 of every path:  yhat[b, t-t0, k, a] = de-standardised predicted log-return
 * of asset a, k+1 days ahead of embedded row (row0 + t) — the loop at backtest.py:99-121 for all t at once.
 * z: standardised series [B,T,ld_z] fp32 (kmpc_standardize output); embedded row i reads days i..i+d-1.
 * mean/std as in kmpc_current_returns. */
int kmpc_forecast(kmpc_handle* h, const kmpc_model* m, const float* z, int ld_z, const double* mean,
                  const double* std, int stats_per_path, int B, int T, int row0, int t0, int t1, int H,
                  float* yhat, void* stream);

/* model.encode on explicit embedded rows obs[M,obs] -> latent[M,Z]  (parity hook, model.py:765, 837) */
int kmpc_encode(kmpc_handle* h, const kmpc_model* m, const float* obs, int M, float* latent, void* stream);
/* KoopmanMachine.step_latent (model.py:311-321, 787-797): out[M,Z] = norm(z[M,Z] @ kmat) */
int kmpc_step_latent(kmpc_handle* h, const kmpc_model* m, const float* z, int M, float* out, void* stream);
/* KoopmanMachine.decode (model.py:768-777, 839-850): out[M,obs] */
int kmpc_decode(kmpc_handle* h, const kmpc_model* m, const float* z, int M, float* out, void* stream);
/* rollout on explicit embedded rows: pred[M,H,obs_cols] standardised decoder output, first obs_cols columns */
int kmpc_rollout(kmpc_handle* h, const kmpc_model* m, const float* obs, int M, int H, int obs_cols, float* pred,
                 void* stream);

/* diagnostics: choose the GEMM kernel family process-wide (1 = tcgen05 3xTF32 where eligible [default],
 * 0 = fp32 SIMT everywhere) and run one bare GEMM C[M,Nout] = A[M,K] . W[Nout,K]^T through a chosen kernel
 * (mode 0 = SIMT, 1 = tcgen05 3xTF32, 2 = tcgen05 fp16 pairs; KMPC_E_CUDA with "not eligible" if the shape cannot use it). */
int kmpc_set_gemm_mode(int use_tensor_cores);
/* 1 [default]: when the latent step and the read-out are linear (GenericKM with NORM_FN 'id' and a one-layer
 * decoder, or LISTAKM) kmpc_forecast evaluates all H horizons with ONE GEMM against the pre-multiplied matrices
 * D_N (K^T)^(k+1) (built in fp64 at first use); 0: always step z <- z K and decode H times like backtest.py:107-121.
 * Same forecasts within fp32 rounding (tests: 1e-5 relative, norm-wise). */
int kmpc_set_forecast_fold(int on);
/* 1 [default]: the GenericKM encoder + folded read-out of kmpc_forecast run on the fp16-pair tensor-core kernel
 * (an fp32 value travels as fp16(x) and fp16((x - hi) * 2^11): 22 significant bits like the 3xTF32 pair at half the
 * operand bytes and twice the MMA rate).  Its epilogues raise a device flag when a value leaves the fp16 range
 * (|x| > 65504); the 3xTF32 chain is queued behind it gated on that flag (its kernels exit at once otherwise), so the
 * call stays asynchronous on `stream`.  0: always the 3xTF32 chain.  2: as 1, on the CTA-pair instantiation of the kernel
 * (tcgen05 cta_group::2, 256 x 128 tiles: each CTA stages half of the weight tile). */
int kmpc_set_gemm_fp16_pairs(int on);
/* diagnostics / tuning: rows per pass of the fp16-pair forecast chain (rounded down to whole paths; default: see
 * forecast.cu).  Smaller passes keep the layer activations L2-resident between the GEMMs of the chain. */
int kmpc_set_forecast_chunk_rows(int rows);
/* 1 [default]: the first layer of the fp16-pair chain reads a materialised embedding of each pass (fp16 pair, rows padded
 * to 128 bytes, written by one gather kernel per pass: the time-delay embedding of data_finance.py:262-300 for the rows
 * of the pass); 0: it reads the delay windows in place through a 3-D tensor map (no copy, but unaligned operand rows
 * and K = delay * (n_assets rounded up to 8)).  Same forecasts up to the summation order of the first layer. */
int kmpc_set_forecast_embedding(int materialise);
int kmpc_debug_gemm(kmpc_handle* h, const float* A, const float* W, int M, int Nout, int K, float* C, int mode);

/* ---------------------------------------------------------------------------------------------
 * MPC — replaces mpc.solve_mpc_log_utility (mpc.py:27-117)
 * ------------------------------------------------------------------------------------------- */

/* P independent problems.  yhat [P,H,N] fp32 predicted log-returns (R = fp32 exp as mpc.py:55 on the fp32
 * array the strategy passes), or pass yhat = NULL and yhat64 [P,H,N] fp64 log-returns (R = fp64 exp, what
 * mpc.py:55 does for a float64 array, e.g. reference tests/test_mpc.py).  lam/tau: [P] arrays or NULL to use the scalars
 * (MPCConfig.cost_coeff / max_turnover; tau <= 0 = no cap, mpc.py:94).  Outputs (any may be NULL except w_out):
 * w_out [P,H,N] fp64, obj [P] maximised objective (NaN on fallback), kkt [P,3] = primal residual, dual
 * residual, complementarity gap, status [P] KMPC_STATUS_*, iters [P]. */
int kmpc_mpc_solve(kmpc_handle* h, const float* yhat, const double* yhat64, const double* w_cur,
                   const double* lam, const double* tau, double lam0, double tau0, int allow_short, int P, int H,
                   int N, double* w_out, double* obj, double* kkt, int32_t* status, int32_t* iters, void* stream);

/* Same, HOST buffers, one call = H2D + solve + D2H + sync (the P=1 drop-in used by the Python shim). */
int kmpc_mpc_solve_host(kmpc_handle* h, const void* yhat_host, int yhat_is_f64, const double* w_cur_host,
                        double lam0, double tau0, int allow_short, int P, int H, int N, double* w_out_host,
                        double* obj_host, double* kkt_host, int32_t* status_host, int32_t* iters_host);

/* ---------------------------------------------------------------------------------------------
 * Mean-variance MPC — replaces mpc.solve_mpc_mean_variance (mpc.py:119-184; used by MarkowitzStrategy,
 * baselines.py:24-106, with H = 1)
 * ------------------------------------------------------------------------------------------- */

/* P independent problems:  max sum_t [ w_t.mu_t - gamma w_t' Sigma w_t ] - lam sum_t ||w_t - w_{t-1}||_1,
 * sum(w_t) = 1, w_t >= 0 unless allow_short, no turnover cap.  mu [P,H,N] fp64, sigma [N,N] shared or [P,N,N]
 * (sigma_per_problem != 0), w_cur [P,N].  Outputs as kmpc_mpc_solve; on failure w_out = tile(w_cur), obj = NaN
 * (mpc.py:179-180).  H * N <= 160 runs one warp per problem with the dense Newton matrix in shared memory; up to
 * H * N = 1280 (H <= 8) one block per problem with the matrix in a handle-owned global workspace; beyond that
 * KMPC_E_UNSUPPORTED. */
int kmpc_mv_supported(int H, int N);
int kmpc_mpc_mean_variance(kmpc_handle* h, const double* mu, const double* sigma, int sigma_per_problem, const double* w_cur,
                           double gamma, double lam, int allow_short, int P, int H, int N, double* w_out, double* obj,
                           double* kkt, int32_t* status, int32_t* iters, void* stream);
/* Same, one problem, HOST buffers (the drop-in used by the Python shim). */
int kmpc_mpc_mean_variance_host(kmpc_handle* h, const double* mu_host, const double* sigma_host, const double* w_cur_host,
                                double gamma, double lam, int allow_short, int H, int N, double* w_out_host, double* obj_host,
                                double* kkt_host, int32_t* status_host, int32_t* iters_host);

/* ---------------------------------------------------------------------------------------------
 * Backtest — replaces run_backtest's step loop + calculate_metrics (backtest.py:133-249)
 * ------------------------------------------------------------------------------------------- */

typedef struct kmpc_backtest_desc {
  int B, N, H;
  int rows;            /* rows of the test split (len(dataset)+sequence_length) */
  int n_steps;         /* len(test_dataset) - horizon (backtest.py:150) */
  int rebalance_freq;  /* backtest.py:163 */
  int allow_short;
  const float* yhat;          /* forecasts [n_yhat_sets, n_steps, H, N] */
  const int32_t* yhat_index;  /* [B] which forecast set a backtest uses, or NULL = b */
  const float* realized;      /* [n_paths, rows, N] kmpc_current_returns output */
  const int32_t* realized_index; /* [B] or NULL = b */
  const double* lam;  const double* tau;  const double* cost_coeff;  const double* capital; /* [B] or NULL */
  double lam0, tau0;          /* MPCConfig.cost_coeff, MPCConfig.max_turnover */
  double cost_coeff0;         /* BacktestConfig.cost_coeff */
  double capital0;            /* BacktestConfig.initial_capital */
  double* history;            /* [B, ceil(n_steps/rebalance_freq), 4] value, return, turnover, cost; or NULL */
  double* metrics;            /* [B,5] Sharpe, Max Drawdown, Avg Turnover, Final Value, Total Return */
  int64_t* solve_stats;       /* [B,4] #optimal, #inaccurate, #fallback, total solver iterations; or NULL */
  double* final_weights;      /* [B,N] or NULL */
} kmpc_backtest_desc;

/* Asynchronous on `stream`.  One persistent kernel, or (KMPC_PARAM_ACTIVE_SET, default) a fixed sequence of five: the dense
 * start of every backtest, the reduced solves once its portfolio has concentrated, and — for backtests the reduced-solve
 * kernel had to give back — a full-width pass, a second reduced-solve pass and a final full-width pass; they hand
 * backtests to each other through a handle-owned state buffer ([B, N + 16] doubles), without host synchronisation.  `solve_stats` counts every Newton step a decision took, re-solves on a grown active set included. */
int kmpc_backtest_run(kmpc_handle* h, const kmpc_backtest_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KMPC_H_ */
