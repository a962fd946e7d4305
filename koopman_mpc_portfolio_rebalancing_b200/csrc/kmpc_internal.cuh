// Internal declarations shared by the .cu files of libkmpc (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "per_device.cuh"
#include "mpc_common.cuh"

namespace kmpc {

struct MpcSolveArgs {
  const float* yhat;        // [P,H,N] predicted log-returns (fp32, as the strategy passes them) or null
  const double* yhat64;     // [P,H,N] fp64 log-returns (R = exp in fp64) when yhat == null
  const double* w_cur;      // [P,N]
  const double* lam;        // [P] or null -> lam0
  const double* tau;        // [P] or null -> tau0
  double lam0, tau0;
  int allow_short, P, N;
  double* w_out;            // [P,H,N]
  double* obj;              // [P]  (NaN when the fallback was taken)
  double* kkt;              // [P,3] primal residual, dual residual, complementarity gap
  int* status;              // [P]
  int* iters;               // [P]
  int* fix_flag;            // device int (handle-owned scratch): 1 = every problem has lam > 0 and tau > 0
  IpmOptions opt;
};

struct BacktestArgs {
  const float* yhat;        // forecasts, row t of backtest b at yhat + index(b)*yhat_stride + t*H*N
  const float* realized;    // de-standardised log-returns of every test row, [.., rows, N]
  const int* yhat_index;    // [B] or null (identity)
  const int* realized_index;
  long long yhat_stride, realized_stride;   // elements
  int rows, n_steps, rebalance_freq, n_hist;
  const double* lam; const double* tau; const double* cost_coeff; const double* capital;   // [B] or null
  double lam0, tau0, cost_coeff0, capital0;
  int allow_short, B, N;
  double* history;          // [B, n_hist, 4] or null: portfolio_value, return, turnover, cost
  double* metrics;          // [B,5] Sharpe, MaxDD, AvgTurnover, FinalValue, TotalReturn
  long long* solve_stats;   // [B,4] or null: #optimal, #inaccurate, #fallback, total IPM iterations
  double* final_weights;    // [B,N] or null
  int* work_counter;        // device int, zeroed before launch
  int* fix_flag;            // device int (handle-owned scratch), see MpcSolveArgs
  IpmOptions opt;
  // active-set pipeline (mpc_lane_kernels.cuh, backtest_active_kernel): three launches share per-backtest state
  int phase;                // 0: whole backtests (no state); 1: dense start, hands over; 3: resumes suspended backtests and hands
                            // them over again; 2: resumes suspended backtests, to their end
  double* state;            // [B, state_ld]: weights [N], book-keeping, step index; null = pipeline not available
  int* bt_status;           // [B] 0 fresh, 1 ready for the active-set kernel, 2 suspended (needs the full solver), 3 done
  int state_ld;
  int as_hmax;              // phase 1 hands a backtest over at the first decision with at most this many held assets
  int* done_counter;        // device int: backtests that have left the active-set kernel for good (finished or suspended)
  int seg;                  // decisions per work item of the active-set kernel (0: whole backtests)
  int* ready_ring;          // [B] ids of the backtests ready for the active-set kernel (-1: empty cell), FIFO
  int* queue_ctr;           // device ints: [0] head, [1] tail of the ring (monotone counters, position = counter % B)
};

int dispatch_mpc_solve(const MpcSolveArgs& A, int H, int sm_count, cudaStream_t st);
int dispatch_backtest(const BacktestArgs& A, int H, int sm_count, cudaStream_t st);
// 1 if this backtest can take the active-set pipeline (the caller then provides A.state / A.bt_status)
int active_set_eligible(const BacktestArgs& A, int H);

}  // namespace kmpc

namespace kmpc {
int mpc_variant_supported(int H, int N);
int mv_supported(int H, int N);
long long mv_work_doubles(int H, int N);
int mv_blocks(int P, int H, int N, int sm_count);
int launch_mpc_mv(const double* mu, const double* sigma, long long sigma_stride, const double* w_cur, double gamma, double lam,
                  int allow_short, int P, int H, int N, double* w_out, double* obj, double* kkt, int* status, int* iters,
                  double* work, int sm_count, cudaStream_t st);
int launch_standardize(const double* y, const double* mean, const double* sd, int spp, int B, int T, int N, float* out,
                       int ld, int sm_count, cudaStream_t st);
int launch_embed_gather(const float* data, int ld, int B, int T, int N, int d, float* out, int sm_count, cudaStream_t st);
int launch_current_returns(const float* z, int ld, const double* mean, const double* sd, int spp, int B, int T, int N,
                           int d, int row0, int rows, float* out, int sm_count, cudaStream_t st);
}  // namespace kmpc

// Entry points run on the handle's device and leave the caller's current device as they found it (a process that
// drives several GPUs — one handle each — shares the CUDA current-device state with its tensor library).
struct kmpc_device_guard {
  int prev = -1;
  cudaError_t err;
  explicit kmpc_device_guard(int dev) {
    cudaGetDevice(&prev);
    err = (prev == dev) ? cudaSuccess : cudaSetDevice(dev);
    if (prev == dev) prev = -1;
  }
  ~kmpc_device_guard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct kmpc_handle {
  int device;
  int sm_count;
  long long launches;
  int* work_counter;      // device ints: [0], [1], [3] work counters of the backtest launches, [2] structure flag, [4] done counter, [5], [6] head / tail of the ready queue
  double* bt_state;       // per-backtest state of the active-set pipeline, grown on demand
  size_t bt_state_doubles;
  int* bt_status;
  size_t bt_status_n;
  void* scratch;          // device workspace (forecast activations), grown on demand
  size_t scratch_bytes;
  double* mv_work;        // device workspace of the block-wide mean-variance kernel, grown on demand
  size_t mv_work_doubles;
  float* stats32;         // device [stats32_cap] fp32 copies of (std, mean) for the forecast epilogue
  int stats32_cap;
  kmpc::IpmOptions ipm;   // solver options of this handle (kmpc_set_solver_param)
};
