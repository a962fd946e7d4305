// extern "C" surface of libkmpc (include/kmpc.h).  Thin: argument checks, launches, error strings.
#include <stdio.h>
#include <string.h>
#include <stdarg.h>
#include "../../include/kmpc.h"
#include "kmpc_internal.cuh"

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int kmpc_fail_cuda(cudaError_t e, const char* what) {
  return fail(KMPC_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
}
#define CK(call)                                                                  \
  do {                                                                            \
    cudaError_t e_ = (call);                                                      \
    if (e_ != cudaSuccess) return kmpc_fail_cuda(e_, #call);                      \
  } while (0)

extern "C" {

int kmpc_version(void) { return KMPC_VERSION; }
const char* kmpc_last_error(void) { return g_err; }

int kmpc_create(int device, kmpc_handle** out) {
  if (!out) return fail(KMPC_E_INVALID, "kmpc_create: out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(KMPC_E_CUDA, "kmpc_create: no CUDA device (%s); libkmpc has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(KMPC_E_INVALID, "kmpc_create: device %d out of range [0,%d)", device, n);
  kmpc_device_guard dev_guard_(device);
  CK(dev_guard_.err);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(KMPC_E_UNSUPPORTED, "kmpc_create: device %d is sm_%d%d; libkmpc is built for sm_100a only", device,
                prop.major, prop.minor);
  kmpc_handle* h = new kmpc_handle();
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->launches = 0;
  h->scratch = nullptr;
  h->scratch_bytes = 0;
  h->stats32 = nullptr;
  h->stats32_cap = 0;
  h->mv_work = nullptr;
  h->mv_work_doubles = 0;
  h->bt_state = nullptr; h->bt_state_doubles = 0; h->bt_status = nullptr; h->bt_status_n = 0;
  h->ipm = kmpc::default_ipm_options();
  e = cudaMalloc(&h->work_counter, 8 * sizeof(int));      // [0], [1], [3] backtest work counters, [2] structure-flag scratch, [4] done counter
  if (e != cudaSuccess) { delete h; return kmpc_fail_cuda(e, "cudaMalloc(work_counter)"); }
  *out = h;
  return KMPC_OK;
}

int kmpc_destroy(kmpc_handle* h) {
  if (!h) return KMPC_OK;
  kmpc_device_guard dev_guard_(h->device);
  cudaFree(h->work_counter);
  if (h->bt_state) cudaFree(h->bt_state);
  if (h->bt_status) cudaFree(h->bt_status);
  if (h->scratch) cudaFree(h->scratch);
  if (h->stats32) cudaFree(h->stats32);
  if (h->mv_work) cudaFree(h->mv_work);
  delete h;
  return KMPC_OK;
}

int64_t kmpc_launch_count(const kmpc_handle* h) { return h ? h->launches : 0; }
int kmpc_mpc_supported(int H, int N) { return kmpc::mpc_variant_supported(H, N); }
int kmpc_set_solver_param(kmpc_handle* h, int which, double value) {
  if (!h) return fail(KMPC_E_INVALID, "kmpc_set_solver_param: NULL handle");
  if (!(value == value)) return fail(KMPC_E_INVALID, "kmpc_set_solver_param: NaN");
  switch (which) {
    case KMPC_PARAM_STEP_FRAC:
      if (!(value > 0.0 && value < 1.0)) return fail(KMPC_E_INVALID, "kmpc_set_solver_param: step_frac must be in (0,1)");
      h->ipm.step_frac = value; break;
    case KMPC_PARAM_DUAL_INIT:
      if (value < 0.0) return fail(KMPC_E_INVALID, "kmpc_set_solver_param: dual_init must be >= 0");
      h->ipm.dual_init = value; break;
    case KMPC_PARAM_MAX_ITER:
      if (value < 1.0 || value > 1000.0) return fail(KMPC_E_INVALID, "kmpc_set_solver_param: max_iter must be in [1,1000]");
      h->ipm.max_iter = (int)value; break;
    case KMPC_PARAM_TOL:
      if (!(value >= 1e-13 && value <= 1e-6)) return fail(KMPC_E_INVALID, "kmpc_set_solver_param: tol must be in [1e-13, 1e-6]");
      h->ipm.tol = value; break;
    case KMPC_PARAM_SECOND_ATTEMPT:
      h->ipm.second_attempt = (value != 0.0) ? 1 : 0; break;
    case KMPC_PARAM_CLIP_FIRST_TRADE:
      h->ipm.clip_first_trade = (value != 0.0) ? 1 : 0; break;
    case KMPC_PARAM_ACTIVE_SET:
      if (value >= 8.0) { h->ipm.active_set = 1; h->ipm.active_seg = (int)value; }            // tuning: decisions per work item
      else { h->ipm.active_set = (value == 2.0) ? 2 : ((value != 0.0) ? 1 : 0); h->ipm.active_seg = (value == 3.0) ? 0 : 32; }
      break;
    case KMPC_PARAM_RESET:
      h->ipm = kmpc::default_ipm_options(); break;
    default:
      return fail(KMPC_E_INVALID, "kmpc_set_solver_param: unknown parameter %d", which);
  }
  return KMPC_OK;
}

int kmpc_standardize(kmpc_handle* h, const double* logret, const double* mean, const double* std, int stats_per_path,
                     int B, int T, int N, float* out, int ld_out, void* stream) {
  if (!h || !logret || !mean || !std || !out) return fail(KMPC_E_INVALID, "kmpc_standardize: NULL argument");
  if (B <= 0 || T <= 0 || N <= 0 || ld_out < N) return fail(KMPC_E_INVALID, "kmpc_standardize: bad shape");
  kmpc_device_guard dev_guard_(h->device);
  CK(dev_guard_.err);
  int rc = kmpc::launch_standardize(logret, mean, std, stats_per_path, B, T, N, out, ld_out, h->sm_count, (cudaStream_t)stream);
  h->launches++;
  if (rc) return kmpc_fail_cuda((cudaError_t)rc, "standardize_kernel");
  return KMPC_OK;
}

int kmpc_embed_gather(kmpc_handle* h, const float* data, int ld_in, int B, int T, int N, int d, float* out, void* stream) {
  if (!h || !data || !out) return fail(KMPC_E_INVALID, "kmpc_embed_gather: NULL argument");
  if (T < d) return fail(KMPC_E_INVALID, "Time series length %d < embedding_dim %d", T, d);
  if (B <= 0 || N <= 0 || d <= 0 || ld_in < N) return fail(KMPC_E_INVALID, "kmpc_embed_gather: bad shape");
  kmpc_device_guard dev_guard_(h->device);
  CK(dev_guard_.err);
  int rc = kmpc::launch_embed_gather(data, ld_in, B, T, N, d, out, h->sm_count, (cudaStream_t)stream);
  h->launches++;
  if (rc) return kmpc_fail_cuda((cudaError_t)rc, "embed_gather_kernel");
  return KMPC_OK;
}

int kmpc_embed_index_host(int T, int N, int d, int32_t* idx) {
  if (!idx || N <= 0 || d <= 0) return fail(KMPC_E_INVALID, "kmpc_embed_index_host: bad argument");
  if (T < d) return fail(KMPC_E_INVALID, "Time series length %d < embedding_dim %d", T, d);
  const int rows = T - d + 1;
  for (int i = 0; i < rows; ++i)
    for (int j = 0; j < d; ++j)
      for (int a = 0; a < N; ++a) idx[((size_t)i * d + j) * N + a] = (i + d - 1 - j) * N + a;
  return KMPC_OK;
}

int kmpc_current_returns(kmpc_handle* h, const float* z, int ld_z, const double* mean, const double* std,
                         int stats_per_path, int B, int T, int N, int d, int row0, int rows, float* out, void* stream) {
  if (!h || !z || !mean || !std || !out) return fail(KMPC_E_INVALID, "kmpc_current_returns: NULL argument");
  if (B <= 0 || rows <= 0 || N <= 0 || d <= 0 || row0 < 0 || row0 + rows + d - 1 > T || ld_z < N)
    return fail(KMPC_E_INVALID, "kmpc_current_returns: bad shape (row0=%d rows=%d d=%d T=%d)", row0, rows, d, T);
  kmpc_device_guard dev_guard_(h->device);
  CK(dev_guard_.err);
  int rc = kmpc::launch_current_returns(z, ld_z, mean, std, stats_per_path, B, T, N, d, row0, rows, out, h->sm_count,
                                        (cudaStream_t)stream);
  h->launches++;
  if (rc) return kmpc_fail_cuda((cudaError_t)rc, "current_returns_kernel");
  return KMPC_OK;
}

int kmpc_mpc_solve(kmpc_handle* h, const float* yhat, const double* yhat64, const double* w_cur, const double* lam,
                   const double* tau, double lam0, double tau0, int allow_short, int P, int H, int N, double* w_out,
                   double* obj, double* kkt, int32_t* status, int32_t* iters, void* stream) {
  if (!h || (!yhat && !yhat64) || !w_cur || !w_out) return fail(KMPC_E_INVALID, "kmpc_mpc_solve: NULL argument");
  if (P < 0 || H <= 0 || N <= 0) return fail(KMPC_E_INVALID, "kmpc_mpc_solve: bad shape P=%d H=%d N=%d", P, H, N);
  if (P == 0) return KMPC_OK;
  if (!kmpc::mpc_variant_supported(H, N))
    return fail(KMPC_E_UNSUPPORTED, "kmpc_mpc_solve: no compiled kernel variant for H=%d N=%d", H, N);
  kmpc_device_guard dev_guard_(h->device);
  CK(dev_guard_.err);
  kmpc::MpcSolveArgs A;
  A.yhat = yhat; A.yhat64 = yhat64; A.w_cur = w_cur; A.lam = lam; A.tau = tau; A.lam0 = lam0; A.tau0 = tau0;
  A.allow_short = allow_short; A.P = P; A.N = N; A.w_out = w_out; A.obj = obj; A.kkt = kkt; A.status = status;
  A.iters = iters; A.fix_flag = h->work_counter + 2; A.opt = h->ipm;
  int rc = kmpc::dispatch_mpc_solve(A, H, h->sm_count, (cudaStream_t)stream);
  h->launches++;
  if (rc == -2) return fail(KMPC_E_UNSUPPORTED, "kmpc_mpc_solve: unsupported shape");
  if (rc) return kmpc_fail_cuda((cudaError_t)rc, "mpc_solve_kernel");
  return KMPC_OK;
}

int kmpc_mpc_solve_host(kmpc_handle* h, const void* yhat_host, int yhat_is_f64, const double* w_cur_host, double lam0,
                        double tau0, int allow_short, int P, int H, int N, double* w_out_host, double* obj_host,
                        double* kkt_host, int32_t* status_host, int32_t* iters_host) {
  if (!h || !yhat_host || !w_cur_host || !w_out_host) return fail(KMPC_E_INVALID, "kmpc_mpc_solve_host: NULL argument");
  if (P <= 0 || H <= 0 || N <= 0) return fail(KMPC_E_INVALID, "kmpc_mpc_solve_host: bad shape");
  kmpc_device_guard dev_guard_(h->device);
  CK(dev_guard_.err);
  const size_t nw = (size_t)P * H * N;
  const size_t ybytes = nw * (yhat_is_f64 ? sizeof(double) : sizeof(float));
  const size_t bytes = nw * sizeof(double) + (size_t)P * N * sizeof(double) + nw * sizeof(double) +
                       (size_t)P * (4 * sizeof(double) + 2 * sizeof(int32_t)) + 64;
  char* buf = nullptr;
  CK(cudaMalloc(&buf, bytes));
  double* d_w = (double*)buf;                       // keep 8-byte quantities first
  double* d_wc = d_w + nw;
  double* d_obj = d_wc + (size_t)P * N;
  double* d_kkt = d_obj + P;
  double* d_y = d_kkt + 3 * (size_t)P;               // fp32 or fp64 log-returns
  int32_t* d_st = (int32_t*)(d_y + nw);
  int32_t* d_it = d_st + P;
  int rc = KMPC_OK;
  cudaError_t e;
  if ((e = cudaMemcpy(d_y, yhat_host, ybytes, cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(d_wc, w_cur_host, (size_t)P * N * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) {
    cudaFree(buf);
    return kmpc_fail_cuda(e, "H2D");
  }
  rc = kmpc_mpc_solve(h, yhat_is_f64 ? nullptr : (const float*)d_y, yhat_is_f64 ? d_y : nullptr, d_wc, nullptr, nullptr, lam0, tau0, allow_short, P, H, N, d_w, d_obj, d_kkt, d_st,
                      d_it, nullptr);
  if (rc == KMPC_OK) {
    e = cudaMemcpy(w_out_host, d_w, nw * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && obj_host) e = cudaMemcpy(obj_host, d_obj, P * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && kkt_host) e = cudaMemcpy(kkt_host, d_kkt, 3 * (size_t)P * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && status_host) e = cudaMemcpy(status_host, d_st, P * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && iters_host) e = cudaMemcpy(iters_host, d_it, P * sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = kmpc_fail_cuda(e, "D2H");
  }
  cudaFree(buf);
  return rc;
}

int kmpc_mv_supported(int H, int N) { return kmpc::mv_supported(H, N); }

int kmpc_mpc_mean_variance(kmpc_handle* h, const double* mu, const double* sigma, int sigma_per_problem, const double* w_cur,
                           double gamma, double lam, int allow_short, int P, int H, int N, double* w_out, double* obj,
                           double* kkt, int32_t* status, int32_t* iters, void* stream) {
  if (!h || !mu || !sigma || !w_cur || !w_out) return fail(KMPC_E_INVALID, "kmpc_mpc_mean_variance: NULL argument");
  if (P <= 0 || H <= 0 || N <= 0 || !(gamma >= 0.0) || !(lam >= 0.0)) return fail(KMPC_E_INVALID, "kmpc_mpc_mean_variance: bad argument");
  if (!kmpc::mv_supported(H, N))
    return fail(KMPC_E_UNSUPPORTED, "kmpc_mpc_mean_variance: H = %d, H*N = %d outside the compiled range (H <= 8, H*N <= 1280)", H, H * N);
  kmpc_device_guard dev_guard_(h->device);
  CK(dev_guard_.err);
  const size_t need = (size_t)kmpc::mv_work_doubles(H, N) * (size_t)kmpc::mv_blocks(P, H, N, h->sm_count);
  if (need > h->mv_work_doubles) {                // large problems: global workspace for the dense Newton matrix
    if (h->mv_work) { CK(cudaStreamSynchronize((cudaStream_t)stream)); cudaFree(h->mv_work); h->mv_work = nullptr; h->mv_work_doubles = 0; }
    CK(cudaMalloc(&h->mv_work, need * sizeof(double)));
    h->mv_work_doubles = need;
  }
  int rc = kmpc::launch_mpc_mv(mu, sigma, sigma_per_problem ? (long long)N * N : 0, w_cur, gamma, lam, allow_short, P, H, N, w_out,
                               obj, kkt, status, iters, h->mv_work, h->sm_count, (cudaStream_t)stream);
  h->launches++;
  if (rc == -2) return fail(KMPC_E_UNSUPPORTED, "kmpc_mpc_mean_variance: unsupported shape");
  if (rc) return kmpc_fail_cuda((cudaError_t)rc, "mpc_mv_kernel");
  return KMPC_OK;
}

int kmpc_mpc_mean_variance_host(kmpc_handle* h, const double* mu_host, const double* sigma_host, const double* w_cur_host,
                                double gamma, double lam, int allow_short, int H, int N, double* w_out_host, double* obj_host,
                                double* kkt_host, int32_t* status_host, int32_t* iters_host) {
  if (!h || !mu_host || !sigma_host || !w_cur_host || !w_out_host) return fail(KMPC_E_INVALID, "kmpc_mpc_mean_variance_host: NULL argument");
  if (H <= 0 || N <= 0) return fail(KMPC_E_INVALID, "kmpc_mpc_mean_variance_host: bad shape");
  kmpc_device_guard dev_guard_(h->device);
  CK(dev_guard_.err);
  const size_t nw = (size_t)H * N;
  const size_t doubles = nw + (size_t)N * N + N + nw + 1 + 3;
  char* buf = nullptr;
  CK(cudaMalloc(&buf, doubles * sizeof(double) + 2 * sizeof(int32_t) + 64));
  double* d_mu = (double*)buf; double* d_sig = d_mu + nw; double* d_wc = d_sig + (size_t)N * N; double* d_w = d_wc + N;
  double* d_obj = d_w + nw; double* d_kkt = d_obj + 1; int32_t* d_st = (int32_t*)(d_kkt + 3); int32_t* d_it = d_st + 1;
  cudaError_t e;
  if ((e = cudaMemcpy(d_mu, mu_host, nw * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(d_sig, sigma_host, (size_t)N * N * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
      (e = cudaMemcpy(d_wc, w_cur_host, (size_t)N * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) {
    cudaFree(buf);
    return kmpc_fail_cuda(e, "H2D");
  }
  int rc = kmpc_mpc_mean_variance(h, d_mu, d_sig, 0, d_wc, gamma, lam, allow_short, 1, H, N, d_w, d_obj, d_kkt, d_st, d_it, nullptr);
  if (rc == KMPC_OK) {
    e = cudaMemcpy(w_out_host, d_w, nw * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && obj_host) e = cudaMemcpy(obj_host, d_obj, sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && kkt_host) e = cudaMemcpy(kkt_host, d_kkt, 3 * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && status_host) e = cudaMemcpy(status_host, d_st, sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && iters_host) e = cudaMemcpy(iters_host, d_it, sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = kmpc_fail_cuda(e, "D2H");
  }
  cudaFree(buf);
  return rc;
}

int kmpc_backtest_run(kmpc_handle* h, const kmpc_backtest_desc* D, void* stream) {
  if (!h || !D) return fail(KMPC_E_INVALID, "kmpc_backtest_run: NULL argument");
  if (!D->yhat || !D->realized || !D->metrics) return fail(KMPC_E_INVALID, "kmpc_backtest_run: yhat/realized/metrics NULL");
  if (D->B <= 0 || D->N <= 0 || D->H <= 0 || D->rows <= 0 || D->n_steps < 0 || D->rebalance_freq <= 0)
    return fail(KMPC_E_INVALID, "kmpc_backtest_run: bad shape");
  if (!kmpc::mpc_variant_supported(D->H, D->N))
    return fail(KMPC_E_UNSUPPORTED, "kmpc_backtest_run: no compiled kernel variant for H=%d N=%d", D->H, D->N);
  kmpc_device_guard dev_guard_(h->device);
  CK(dev_guard_.err);
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(h->work_counter, 0, 8 * sizeof(int), st));
  kmpc::BacktestArgs A;
  A.yhat = D->yhat; A.realized = D->realized; A.yhat_index = D->yhat_index; A.realized_index = D->realized_index;
  A.yhat_stride = (long long)D->n_steps * D->H * D->N;
  A.realized_stride = (long long)D->rows * D->N;
  A.rows = D->rows; A.n_steps = D->n_steps; A.rebalance_freq = D->rebalance_freq;
  A.n_hist = (D->n_steps + D->rebalance_freq - 1) / D->rebalance_freq;
  A.lam = D->lam; A.tau = D->tau; A.cost_coeff = D->cost_coeff; A.capital = D->capital;
  A.lam0 = D->lam0; A.tau0 = D->tau0; A.cost_coeff0 = D->cost_coeff0; A.capital0 = D->capital0;
  A.allow_short = D->allow_short; A.B = D->B; A.N = D->N;
  A.history = D->history; A.metrics = D->metrics; A.solve_stats = (long long*)D->solve_stats;
  A.final_weights = D->final_weights; A.work_counter = h->work_counter; A.fix_flag = h->work_counter + 2; A.opt = h->ipm;
  A.phase = 0; A.state = nullptr; A.bt_status = nullptr; A.state_ld = D->N + 16; A.as_hmax = 24;       // (28 made config-3 backtests overflow the warp twice and finish full-width: 1.9 s instead of 1.0 s)
  A.done_counter = h->work_counter + 4; A.seg = h->ipm.active_seg;
  A.ready_ring = nullptr; A.queue_ctr = h->work_counter + 5;
  if (kmpc::active_set_eligible(A, D->H)) {
    // per-backtest state of the three-launch pipeline (dense start -> reduced solves -> stragglers)
    const size_t need = (size_t)D->B * A.state_ld;
    if (h->bt_state_doubles < need) {
      if (h->bt_state) cudaFree(h->bt_state);
      h->bt_state = nullptr; h->bt_state_doubles = 0;
      CK(cudaMalloc(&h->bt_state, need * sizeof(double)));
      h->bt_state_doubles = need;
    }
    if (h->bt_status_n < (size_t)D->B) {
      if (h->bt_status) cudaFree(h->bt_status);
      h->bt_status = nullptr; h->bt_status_n = 0;
      CK(cudaMalloc(&h->bt_status, (size_t)2 * D->B * sizeof(int)));          // status [B] | ready ring [B]
      h->bt_status_n = D->B;
    }
    CK(cudaMemsetAsync(h->bt_status, 0, (size_t)D->B * sizeof(int), st));
    CK(cudaMemsetAsync(h->bt_status + h->bt_status_n, 0xff, (size_t)D->B * sizeof(int), st));   // -1: empty cells
    A.state = h->bt_state; A.bt_status = h->bt_status; A.ready_ring = h->bt_status + h->bt_status_n;
  }
  int rc = kmpc::dispatch_backtest(A, D->H, h->sm_count, st);
  h->launches += A.state ? 5 : 1;                     // dense start, reduced solves, second chance (2), stragglers
  if (rc == -2) return fail(KMPC_E_UNSUPPORTED, "kmpc_backtest_run: unsupported shape");
  if (rc) return kmpc_fail_cuda((cudaError_t)rc, "backtest_kernel");
  return KMPC_OK;
}

}  // extern "C"
