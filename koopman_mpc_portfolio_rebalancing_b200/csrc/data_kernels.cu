// Data side of the hot path: standardise, time-delay embedding gather, current-return extraction.
// Replaces data_finance.py:243-300 and :717-742.  All three are HBM-bound streaming kernels:
// coalesced along the asset/column axis, grid sized in multiples of the SM count (grid-stride loops).
#include "kmpc_internal.cuh"

namespace kmpc {

// out[b,t,a] = (float)((y - mean)/std) in fp64 (bit-exact with numpy), padding columns [N, ld) zeroed.
__global__ void standardize_kernel(const double* __restrict__ y, const double* __restrict__ mean,
                                   const double* __restrict__ sd, int stats_per_path, long long rows_total, int T,
                                   int N, float* __restrict__ out, int ld) {
  const long long total = rows_total * ld;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / ld;
    const int a = (int)(idx - row * ld);
    float v = 0.0f;
    if (a < N) {
      const long long b = stats_per_path ? row / T : 0;
      const double m = mean[b * N + a], s = sd[b * N + a];
      v = __double2float_rn(__ddiv_rn(__dsub_rn(y[row * N + a], m), s));
    }
    out[idx] = v;
  }
}

// out[b,i,j*N+a] = data[b,i+d-1-j,a]   (data_finance.py:290-298)
__global__ void embed_gather_kernel(const float* __restrict__ data, int ld, int T, int N, int d, long long n_out_rows,
                                    int rows_per_path, float* __restrict__ out) {
  const int obs = d * N;
  const long long total = n_out_rows * obs;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / obs;
    const int c = (int)(idx - row * obs);
    const int j = c / N, a = c - j * N;
    const long long b = row / rows_per_path;
    const int i = (int)(row - b * rows_per_path);
    out[idx] = data[(b * T + i + d - 1 - j) * (long long)ld + a];
  }
}

// out[b,r,a] = fl32(fl32(z * std32) + mean32): torch mul then add, two roundings (data_finance.py:740-742)
__global__ void current_returns_kernel(const float* __restrict__ z, int ld, const double* __restrict__ mean,
                                       const double* __restrict__ sd, int stats_per_path, int T, int N, int d, int row0,
                                       int rows, long long total, float* __restrict__ out) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long br = idx / N;
    const int a = (int)(idx - br * N);
    const long long b = br / rows;
    const int r = (int)(br - b * rows);
    const long long sb = stats_per_path ? b : 0;
    const float s32 = __double2float_rn(sd[sb * N + a]), m32 = __double2float_rn(mean[sb * N + a]);
    const float x = z[(b * T + row0 + r + d - 1) * (long long)ld + a];
    out[idx] = __fadd_rn(__fmul_rn(x, s32), m32);
  }
}

static int grid_for(long long total, int sm_count) {
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

int launch_standardize(const double* y, const double* mean, const double* sd, int spp, int B, int T, int N, float* out,
                       int ld, int sm_count, cudaStream_t st) {
  const long long rows = (long long)B * T;
  standardize_kernel<<<grid_for(rows * ld, sm_count), 256, 0, st>>>(y, mean, sd, spp, rows, T, N, out, ld);
  return (int)cudaGetLastError();
}
int launch_embed_gather(const float* data, int ld, int B, int T, int N, int d, float* out, int sm_count, cudaStream_t st) {
  const int rpp = T - d + 1;
  const long long n_rows = (long long)B * rpp;
  embed_gather_kernel<<<grid_for(n_rows * d * N, sm_count), 256, 0, st>>>(data, ld, T, N, d, n_rows, rpp, out);
  return (int)cudaGetLastError();
}
int launch_current_returns(const float* z, int ld, const double* mean, const double* sd, int spp, int B, int T, int N,
                           int d, int row0, int rows, float* out, int sm_count, cudaStream_t st) {
  const long long total = (long long)B * rows * N;
  current_returns_kernel<<<grid_for(total, sm_count), 256, 0, st>>>(z, ld, mean, sd, spp, T, N, d, row0, rows, total, out);
  return (int)cudaGetLastError();
}

}  // namespace kmpc
