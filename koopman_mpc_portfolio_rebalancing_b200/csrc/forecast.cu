// Koopman forecast path: encoder (MLP or LISTA) -> H x (z <- z @ K [norm]) -> decoder -> first N columns ->
// de-standardise.  Replaces model.py (MLPCoder 67-117, LISTA 120-209, GenericKM 701-797, LISTAKM 801-870) on the
// backtest path (backtest.py:85-121) for all rebalancing steps of all backtests at once.
//
// Layout: activations row-major fp32 [rows, width], processed in row chunks small enough to stay L2-resident
// between consecutive layers.  The delay-embedded input is never materialised: row (b,t) of the first GEMM is
// the window z[b, t .. t+d-1, :] of the standardised series (forward time order, row stride ld = pad4(N)), and
// the first-layer weight columns are permuted once at load time to undo the [y_t, y_{t-1}, ...] ordering of
// data_finance.py:290-298.
#include <vector>
#include <stdio.h>
#include "../../include/kmpc.h"
#include "kmpc_internal.cuh"
#include <string.h>
#include "gemm.cuh"

int kmpc_fail_cuda(cudaError_t e, const char* what);

struct kmpc_model {
  kmpc_handle* h;
  int kind, obs, N, d, Z, ld, norm_fn;
  int n_enc, enc_act, enc_last_relu;
  std::vector<int> enc_dims;
  std::vector<float*> enc_w, enc_b, enc_w_lo;
  float* enc_w0_win;           // first encoder / We layer re-laid for the in-place window read [h1, d*ld]
  float* enc_w0_win_lo;
  int n_dec, dec_act;
  std::vector<int> dec_dims;
  std::vector<float*> dec_w, dec_b, dec_w_lo;
  float* kmatT;                // [Z,Z] = kmat^T  (GEMM computes A . W^T)
  float* kmatT_lo;
  int lista_linear, lista_loops;
  float lista_thr;
  float* lista_ST;             // S^T
  float* lista_ST_lo;
  float* lista_wdT;            // [obs, Z]: (dict / ||dict||_row.clamp(1e-4))^T
  float* lista_wdT_lo;
  float* z_lo;                 // residual twin of the standardised series of the current forecast call
  size_t z_lo_cap;
  // folded multi-horizon read-out (linear step + linear decoder): row k*N + a of fold_w = D_N[a,:] . (K^T)^(k+1),
  // so that yhat[:, k, :] = z0 . fold_w[k]^T.  Built on first use for the largest H seen.
  int fold_H;
  float* fold_w; float* fold_w_lo; float* fold_b;
  // fp16-pair copies for gemm_tc16.cu (built on first use): encoder layers (layer 0 re-laid for the in-place window
  // read with row stride ld16 = N rounded up to 8), the folded read-out, the standardised series of the current call
  int ld16, fold16_H;
  std::vector<__half*> w16_hi, w16_lo;
  __half *w16e_hi = nullptr, *w16e_lo = nullptr;   // first layer in the reference's column order, rows padded to Kp (embed16_kernel)
  __half *fold16_hi, *fold16_lo;
  __half *z16_hi, *z16_lo;
  size_t z16_cap;
  int* ovf_flag;              // device int: a value left the fp16 range -> the caller re-runs the TF32 chain
  int* ovf_host;              // pinned host copy
  std::vector<void*> owned;
};

namespace kmpc {

__global__ void transpose_kernel(const float* __restrict__ in, int R, int Cc, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? in[(size_t)r * Cc + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) out[(size_t)c * R + r] = tile[threadIdx.x][i];
  }
}

// Wwin[o, jj*ld + a] = W[o, (d-1-jj)*N + a] for a < N, 0 in the padding columns
__global__ void window_permute_kernel(const float* __restrict__ W, int out_f, int N, int d, int ld, float* __restrict__ Wwin) {
  const long long total = (long long)out_f * d * ld;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int o = (int)(idx / (d * ld));
    const int rem = (int)(idx - (long long)o * d * ld);
    const int jj = rem / ld, a = rem - jj * ld;
    Wwin[idx] = (a < N) ? W[(size_t)o * d * N + (size_t)(d - 1 - jj) * N + a] : 0.f;
  }
}

// Embedded rows of one chunk as an fp16 pair, rows padded to Kp = d*N rounded up to 64 halves (128-byte rows):
//   out[i, j*N + a] = split16(z[b, row_a + t + d-1-j, a]),  global row r0 + i = b * rpp + t,  zero in the padding columns
// (data_finance.py:290-298: newest day first).  The fp16-pair GEMM can read the delay windows in place through a 3-D
// tensor map, but then a row of its A operand starts every ld16 * 2 = 112 bytes (50 assets) and K is d * ld16 = 1120:
// TMA fetches box rows that straddle 128-byte lines (measured on flat GEMMs of this shape: +30 % when the row stride is
// not a multiple of 128 bytes) over 18 k-blocks and two ragged tiles per 246-row path.  Writing the chunk's rows once
// (134 MB, ~30 us) makes layer 1 a flat K = 1024 GEMM like layers 2-3.  One thread = 8 consecutive columns.
__global__ void embed16_kernel(const float* __restrict__ z, int ld, long long T, int rpp, int row_a, int N, int d, int Kp,
                               int r0, int rows, __half* __restrict__ hi, __half* __restrict__ lo,
                               int* __restrict__ overflow) {
  extern __shared__ int col_off[];                   // column -> offset inside the window (or -1: padding)
  for (int c = threadIdx.x; c < Kp; c += blockDim.x) {
    const int j = c / N, a = c - j * N;
    col_off[c] = (j < d) ? (d - 1 - j) * ld + a : -1;
  }
  __syncthreads();
  const int oct = Kp >> 3;                           // 8-column groups per row
  const int tpr = oct < (int)blockDim.x ? oct : (int)blockDim.x;      // threads per row
  const int rpi = (int)blockDim.x / tpr;             // rows per block iteration
  const int sr = (int)threadIdx.x / tpr, lr = (int)threadIdx.x - sr * tpr;
  bool ovf = false;
  if (sr < rpi) {
    for (int i = blockIdx.x * rpi + sr; i < rows; i += gridDim.x * rpi) {
      const int R = r0 + i, b = R / rpp;
      const float* zr = z + ((size_t)b * T + row_a + (R - b * rpp)) * ld;
      for (int o = lr; o < oct; o += tpr) {
        const int c0 = o * 8;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int off = col_off[c0 + e];
          v[e] = (off >= 0) ? zr[off] : 0.0f;
        }
        __align__(16) __half h8[8];
        __align__(16) __half l8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const __half h = __float2half_rn(v[e]);
          h8[e] = h;
          l8[e] = __float2half_rn((v[e] - __half2float(h)) * 2048.0f);
          ovf = ovf || !(fabsf(v[e]) <= 65504.0f);
        }
        *reinterpret_cast<uint4*>(hi + (size_t)i * Kp + c0) = *reinterpret_cast<const uint4*>(h8);
        *reinterpret_cast<uint4*>(lo + (size_t)i * Kp + c0) = *reinterpret_cast<const uint4*>(l8);
      }
    }
  }
  if (ovf && overflow) *overflow = 1;
}

// wdT[c, z] = dict[z, c] / max(||dict[z,:]||_2, 1e-4)   (model.py:848-850); one warp per dictionary row z
__global__ void dict_normalize_transpose_kernel(const float* __restrict__ dict, int Z, int obs, float* __restrict__ wdT) {
  const int z = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (z >= Z) return;
  float s = 0.f;
  for (int c = lane; c < obs; c += 32) { const float v = dict[(size_t)z * obs + c]; s = fmaf(v, v, s); }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float nrm = fmaxf(sqrtf(s), 1e-4f);
  for (int c = lane; c < obs; c += 32) wdT[(size_t)c * Z + z] = dict[(size_t)z * obs + c] / nrm;
}

// z <- z / ||z||_2 per row (GenericKM 'ball' norm, model.py:751-752; no epsilon, like the reference)
__device__ __forceinline__ float lo_of(float v) { return tf32_residual(v); }

__global__ void row_normalize_kernel(float* __restrict__ z, float* __restrict__ zlo, int M, int Z) {
  const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  float s = 0.f;
  for (int c = lane; c < Z; c += 32) { const float v = z[(size_t)m * Z + c]; s = fmaf(v, v, s); }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float nrm = sqrtf(s);
  for (int c = lane; c < Z; c += 32) {
    const float v = z[(size_t)m * Z + c] / nrm;
    z[(size_t)m * Z + c] = v;
    if (zlo) zlo[(size_t)m * Z + c] = lo_of(v);
  }
}

__global__ void shrink_kernel(const float* __restrict__ c, float* __restrict__ z, float* __restrict__ zlo, long long n, float thr) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = epilogue_apply(c[i], EPI_SHRINK, thr);
    z[i] = v;
    zlo[i] = lo_of(v);
  }
}

__global__ void f64_to_f32_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __double2float_rn(in[i]);
}

// one warp per output: out[a, j] = sum_i prev[a, i] * K[j, i]   (fp64 accumulation; prev is fp64, K fp32)
__global__ void fold_step_kernel(const double* __restrict__ prev, const float* __restrict__ kmat, int N, int Z,
                                 double* __restrict__ next, float* __restrict__ w32) {
  const long long o = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= (long long)N * Z) return;
  const int a = (int)(o / Z), j = (int)(o - (long long)a * Z);
  double s = 0.0;
  for (int i = lane; i < Z; i += 32) s = fma(prev[(size_t)a * Z + i], (double)kmat[(size_t)j * Z + i], s);
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) { next[o] = s; w32[o] = __double2float_rn(s); }
}
__global__ void f32_to_f64_kernel(const float* __restrict__ in, double* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = (double)in[i];
}
__global__ void tile_bias_kernel(const float* __restrict__ b, int N, int H, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N * H) out[i] = b[i % N];
}

static int g_fold = 1;
void set_forecast_fold(int on) { g_fold = on ? 1 : 0; }

static int act_to_epi(int act) { return act == KMPC_ACT_RELU ? EPI_RELU : (act == KMPC_ACT_TANH ? EPI_TANH : EPI_GELU); }

struct AView {            // how the rows of the first GEMM are addressed
  const float* A; const float* A_lo; long long group_stride; int rows_per_group; int lda; int K; bool window;
};

static GemmArgs base_args() {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.act = EPI_NONE;
  return g;
}

static int ensure_fold(kmpc_handle* h, kmpc_model* m, int H, cudaStream_t st);

// Runs the whole chain for `M` rows.  out_mode 0: latent z0 -> out [M,Z];  1: yhat de-standardised [M,H,N];
// 2: standardised decoder output, first n_cols columns [M,H,n_cols].
static int run_chain(kmpc_handle* h, const kmpc_model* m, const AView& av, int M, int H, int out_mode, int n_cols,
                     const float* std32, const float* mean32, int stat_rows_per_group, float* out, cudaStream_t st,
                     const int* gate = nullptr) {
  const int Z = m->Z;
  auto gated_args = [&] { GemmArgs g = base_args(); g.gate = gate; return g; };
  int maxw = Z;
  for (int v : m->enc_dims) if (v > maxw) maxw = v;
  for (int v : m->dec_dims) if (v > maxw) maxw = v;
  // Row chunks: 4 ping-pong activation buffers + their residual twins.  The chain is compute bound (activation
  // traffic is ~16 KB per row and layer against ~2 MFLOP), so the chunk is sized for full waves of 128-row tiles,
  // not for L2 residency; window views are chunked in whole paths so that tiles never straddle two paths.
  // A gated (fallback) chain costs one empty launch per kernel when its gate stays shut (2.1 us each, CUPTI timeline):
  // few large passes (32 KB of scratch per row: 4 GB at 131 072 rows) instead of the 500 launches per config-2 step
  // that passes of 8 192 rows made (1.0 ms, 4 % of the forecast stage).
  long long ch = gate ? 131072 : 32768;
  if (av.window && av.rows_per_group < M) {
    ch = (ch / av.rows_per_group) * av.rows_per_group;
    if (ch < av.rows_per_group) ch = av.rows_per_group;
  }
  if (ch > M) ch = M;
  const int CH = (int)ch;
  const size_t need = (size_t)CH * maxw * 8 * sizeof(float) + 256;
  if (h->scratch_bytes < need) {
    if (h->scratch) cudaFree(h->scratch);
    h->scratch = nullptr; h->scratch_bytes = 0;
    cudaError_t e = cudaMalloc(&h->scratch, need);
    if (e != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(forecast scratch)");
    h->scratch_bytes = need;
  }
  float* buf[4];
  float* lob[4];
  for (int i = 0; i < 4; ++i) {
    buf[i] = (float*)h->scratch + (size_t)i * CH * maxw;
    lob[i] = (float*)h->scratch + (size_t)(4 + i) * CH * maxw;
  }
  auto lo_of_buf = [&](const float* p) -> float* {
    for (int i = 0; i < 4; ++i) if (p == buf[i]) return lob[i];
    return nullptr;
  };
  int rc;
  // z_{k+1} = z_k K and a linear read-out compose into one matrix per horizon (model.py:311-321 with NORM_FN 'id',
  // decoder = one Linear, or the LISTAKM dictionary): the H sequential [rows,Z]x[Z,Z] products collapse into one
  // [rows,Z]x[Z,H*N] product.  Algebraically identical, rounding differs at the fp32 level (tests: 1e-5 norm-wise).
  const bool fold = g_fold && out_mode == 1 && H >= 1 &&
                    ((m->kind == KMPC_MODEL_GENERIC && m->norm_fn == KMPC_NORM_ID && m->n_dec == 1) ||
                     (m->kind == KMPC_MODEL_LISTA));
  if (fold && (rc = ensure_fold(h, const_cast<kmpc_model*>(m), H, st))) return rc;
  for (int r0 = 0; r0 < M; r0 += CH) {
    const int rows = (M - r0 < CH) ? (M - r0) : CH;
    // ---------------- encoder ----------------
    const float* x = nullptr; int xld = 0;
    float* zcur = buf[2];
    float* cbuf = buf[3];                 // LISTA pre-activation c
    const bool mlp_enc = (m->kind == KMPC_MODEL_GENERIC) || !m->lista_linear;
    if (mlp_enc) {
      for (int li = 0; li < m->n_enc; ++li) {
        GemmArgs g = gated_args();
        const bool last = (li == m->n_enc - 1);
        if (li == 0) {
          g.A = av.A; g.A_lo = av.A_lo; g.a_group_stride = av.group_stride; g.a_rows_per_group = av.rows_per_group; g.lda = av.lda;
          g.row0 = r0; g.K = av.K;
          g.W = av.window ? m->enc_w0_win : m->enc_w[0]; g.W_lo = av.window ? m->enc_w0_win_lo : m->enc_w_lo[0]; g.ldw = av.K;
        } else {
          g.A = x; g.A_lo = lo_of_buf(x); g.a_group_stride = 0; g.a_rows_per_group = rows; g.lda = xld; g.row0 = 0; g.K = m->enc_dims[li];
          g.W = m->enc_w[li]; g.W_lo = m->enc_w_lo[li]; g.ldw = m->enc_dims[li];
        }
        g.M = rows; g.Nout = m->enc_dims[li + 1]; g.n_store = g.Nout;
        g.bias = m->enc_b[li];
        g.act = last ? (m->enc_last_relu ? EPI_RELU : EPI_NONE) : act_to_epi(m->enc_act);
        float* o = last ? ((m->kind == KMPC_MODEL_LISTA) ? cbuf : zcur) : buf[li & 1];
        g.C = o; g.C_lo = lo_of_buf(o); g.ldc = g.Nout;
        if ((rc = launch_gemm(g, st, &h->launches))) return rc;
        x = o; xld = g.Nout;
      }
    } else {   // LISTA linear encoder: c = x @ We^T
      GemmArgs g = gated_args();
      g.A = av.A; g.A_lo = av.A_lo; g.a_group_stride = av.group_stride; g.a_rows_per_group = av.rows_per_group; g.lda = av.lda; g.row0 = r0;
      g.K = av.K; g.W = av.window ? m->enc_w0_win : m->enc_w[0]; g.W_lo = av.window ? m->enc_w0_win_lo : m->enc_w_lo[0]; g.ldw = av.K;
      g.M = rows; g.Nout = Z; g.n_store = Z; g.C = cbuf; g.ldc = Z;
      if ((rc = launch_gemm(g, st, &h->launches))) return rc;
    }
    if (m->kind == KMPC_MODEL_LISTA) {
      const long long n = (long long)rows * Z;
      int blocks = (int)((n + 255) / 256); if (blocks > h->sm_count * 8) blocks = h->sm_count * 8;
      shrink_kernel<<<blocks, 256, 0, st>>>(cbuf, zcur, lo_of_buf(zcur), n, m->lista_thr); h->launches++;
      float* zalt = buf[0];
      for (int it = 0; it < m->lista_loops; ++it) {          // z = shrink(z @ S + c)
        GemmArgs g = gated_args();
        g.A = zcur; g.A_lo = lo_of_buf(zcur); g.a_rows_per_group = rows; g.lda = Z; g.K = Z; g.W = m->lista_ST; g.W_lo = m->lista_ST_lo; g.ldw = Z;
        g.M = rows; g.Nout = Z; g.n_store = Z; g.addend = cbuf; g.ld_add = Z; g.act = EPI_SHRINK; g.shrink_thr = m->lista_thr;
        g.C = zalt; g.C_lo = lo_of_buf(zalt); g.ldc = Z;
        if ((rc = launch_gemm(g, st, &h->launches))) return rc;
        float* t = zcur; zcur = zalt; zalt = t;
      }
      if (zcur != buf[2]) {   // keep the convention zcur == buf[2] or buf[0]; both are fine below
      }
    } else if (m->norm_fn == KMPC_NORM_BALL) {
      row_normalize_kernel<<<(rows + 7) / 8, 256, 0, st>>>(zcur, lo_of_buf(zcur), rows, Z); h->launches++;
    }
    if (out_mode == 0) {
      cudaError_t e = cudaMemcpyAsync(out + (size_t)r0 * Z, zcur, (size_t)rows * Z * sizeof(float), cudaMemcpyDeviceToDevice, st);
      if (e != cudaSuccess) return kmpc_fail_cuda(e, "copy latent");
      continue;
    }
    // ---------------- folded read-out: all horizons in one GEMM [rows, Z] x [H*N, Z]^T ----------------
    if (fold) {
      GemmArgs d = gated_args();
      d.A = zcur; d.A_lo = lo_of_buf(zcur); d.a_rows_per_group = rows; d.lda = Z; d.K = Z;
      d.W = m->fold_w; d.W_lo = m->fold_w_lo; d.ldw = Z;
      d.M = rows; d.Nout = H * m->N; d.n_store = H * m->N; d.bias = m->fold_b;
      d.C = out + (size_t)r0 * H * m->N; d.ldc = (long long)H * m->N;
      d.std32 = std32; d.mean32 = mean32; d.stat_rows_per_group = stat_rows_per_group; d.stat_ld = m->N; d.stat_row0 = r0;
      d.stat_mod = m->N;
      if ((rc = launch_gemm(d, st, &h->launches))) return rc;
      continue;
    }
    // ---------------- K unroll + decoder ----------------
    float* znext = (zcur == buf[2]) ? buf[0] : buf[2];
    float* hbuf[2] = {buf[1], buf[3]};
    for (int k = 0; k < H; ++k) {
      GemmArgs g = gated_args();
      g.A = zcur; g.A_lo = lo_of_buf(zcur); g.a_rows_per_group = rows; g.lda = Z; g.K = Z; g.W = m->kmatT; g.W_lo = m->kmatT_lo; g.ldw = Z;
      g.M = rows; g.Nout = Z; g.n_store = Z; g.C = znext; g.C_lo = lo_of_buf(znext); g.ldc = Z;
      if ((rc = launch_gemm(g, st, &h->launches))) return rc;
      { float* t = zcur; zcur = znext; znext = t; }
      if (m->kind == KMPC_MODEL_GENERIC && m->norm_fn == KMPC_NORM_BALL) {
        row_normalize_kernel<<<(rows + 7) / 8, 256, 0, st>>>(zcur, lo_of_buf(zcur), rows, Z); h->launches++;
      }
      const int ncol = (out_mode == 1) ? m->N : n_cols;
      float* dst = out + ((size_t)r0 * H + k) * ncol;
      if (m->kind == KMPC_MODEL_GENERIC) {
        const float* xx = zcur; int xl = Z;
        for (int li = 0; li < m->n_dec; ++li) {
          GemmArgs d = gated_args();
          const bool last = (li == m->n_dec - 1);
          d.A = xx; d.A_lo = lo_of_buf(xx); d.a_rows_per_group = rows; d.lda = xl; d.K = m->dec_dims[li]; d.W = m->dec_w[li]; d.W_lo = m->dec_w_lo[li]; d.ldw = m->dec_dims[li];
          d.M = rows; d.bias = m->dec_b[li];
          if (last) {
            d.Nout = ncol; d.n_store = ncol; d.C = dst; d.ldc = (long long)H * ncol;
            if (out_mode == 1) { d.std32 = std32; d.mean32 = mean32; d.stat_rows_per_group = stat_rows_per_group; d.stat_ld = m->N; d.stat_row0 = r0; }
          } else {
            d.Nout = m->dec_dims[li + 1]; d.n_store = d.Nout; d.act = act_to_epi(m->dec_act); d.C = hbuf[li & 1]; d.C_lo = lo_of_buf(d.C); d.ldc = d.Nout;
          }
          if ((rc = launch_gemm(d, st, &h->launches))) return rc;
          xx = d.C; xl = d.Nout;
        }
      } else {
        GemmArgs d = gated_args();
        d.A = zcur; d.A_lo = lo_of_buf(zcur); d.a_rows_per_group = rows; d.lda = Z; d.K = Z; d.W = m->lista_wdT; d.W_lo = m->lista_wdT_lo; d.ldw = Z;
        d.M = rows; d.Nout = ncol; d.n_store = ncol; d.C = dst; d.ldc = (long long)H * ncol;
        if (out_mode == 1) { d.std32 = std32; d.mean32 = mean32; d.stat_rows_per_group = stat_rows_per_group; d.stat_ld = m->N; d.stat_row0 = r0; }
        if ((rc = launch_gemm(d, st, &h->launches))) return rc;
      }
    }
  }
  return 0;
}

}  // namespace kmpc

// ------------------------------------------------------------------------------------------------------------
namespace kmpc {
// fold_w rows [k*N + a] = D_N[a,:] . (K^T)^(k+1) for k < H, built in fp64 and rounded once to fp32 (+ residual twin)
static int ensure_fold(kmpc_handle* h, kmpc_model* m, int H, cudaStream_t st) {
  if (m->fold_w && m->fold_H >= H) return 0;
  const int N = m->N, Z = m->Z;
  const float* DN = (m->kind == KMPC_MODEL_GENERIC) ? m->dec_w[0] : m->lista_wdT;      // [obs, Z], first N rows
  const float* bN = (m->kind == KMPC_MODEL_GENERIC) ? m->dec_b[0] : nullptr;
  float *w = nullptr, *wl = nullptr, *fb = nullptr;
  double* tmp = nullptr;
  cudaError_t e;
  const size_t rows_pad = (size_t)((H * N + 127) / 128) * 128;      // whole 128-row TMA boxes
  if ((e = cudaMalloc(&w, rows_pad * Z * sizeof(float))) != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(fold)");
  if ((e = cudaMalloc(&wl, rows_pad * Z * sizeof(float))) != cudaSuccess) { cudaFree(w); return kmpc_fail_cuda(e, "cudaMalloc(fold)"); }
  if ((e = cudaMalloc(&tmp, (size_t)2 * N * Z * sizeof(double))) != cudaSuccess) { cudaFree(w); cudaFree(wl); return kmpc_fail_cuda(e, "cudaMalloc(fold)"); }
  cudaMemsetAsync(w, 0, rows_pad * Z * sizeof(float), st);
  double* cur = tmp; double* nxt = tmp + (size_t)N * Z;
  f32_to_f64_kernel<<<h->sm_count * 2, 256, 0, st>>>(DN, cur, (long long)N * Z);
  const float* kmat = nullptr;
  // kmatT holds K^T; out[a,j] = sum_i prev[a,i] K[j,i] reads row j of K = column j of K^T: use a transposed copy
  float* kplain = nullptr;
  if ((e = cudaMalloc(&kplain, (size_t)Z * Z * sizeof(float))) != cudaSuccess) { cudaFree(w); cudaFree(wl); cudaFree(tmp); return kmpc_fail_cuda(e, "cudaMalloc(fold)"); }
  {
    dim3 grid((Z + 31) / 32, (Z + 31) / 32), blk(32, 8);
    transpose_kernel<<<grid, blk, 0, st>>>(m->kmatT, Z, Z, kplain);
    kmat = kplain;
  }
  const long long outs = (long long)N * Z;
  for (int k = 0; k < H; ++k) {
    fold_step_kernel<<<(unsigned)((outs + 7) / 8), 256, 0, st>>>(cur, kmat, N, Z, nxt, w + (size_t)k * N * Z);
    double* t = cur; cur = nxt; nxt = t;
  }
  int rc = launch_split_lo(w, wl, (long long)rows_pad * Z, st);
  if (!rc && bN) {
    if ((e = cudaMalloc(&fb, (size_t)H * N * sizeof(float))) != cudaSuccess) rc = (int)e;
    else tile_bias_kernel<<<(H * N + 255) / 256, 256, 0, st>>>(bN, N, H, fb);
  }
  e = cudaStreamSynchronize(st);
  cudaFree(tmp); cudaFree(kplain);
  if (rc || e != cudaSuccess) { cudaFree(w); cudaFree(wl); if (fb) cudaFree(fb); return kmpc_fail_cuda(rc ? (cudaError_t)rc : e, "fold kernels"); }
  if (m->fold_w) { cudaFree(m->fold_w); cudaFree(m->fold_w_lo); if (m->fold_b) cudaFree(m->fold_b); }
  m->fold_w = w; m->fold_w_lo = wl; m->fold_b = fb; m->fold_H = H;
  h->launches += 4 + H;
  return 0;
}
}  // namespace kmpc

namespace kmpc {
static int g_tc16 = 1;
// Rows per pass of the fp16-pair chain (whole paths).  The activations of a pass ping-pong between two fp16-pair buffers
// of CH x width x 4 bytes.
static long long g_chunk_rows = 32768;
void set_forecast_chunk_rows(long long rows) { g_chunk_rows = rows < 128 ? 128 : rows; }
// 1 [default]: the first layer of the fp16-pair chain reads a materialised, 128-byte-aligned embedding of the chunk
// (embed16_kernel); 0: it reads the delay windows in place through the 3-D tensor map
static int g_embed16 = 1;
static int g_alternate = 1;       // tile order alternates from kernel to kernel of a pass (mode 2 of the setter turns it off)
void set_forecast_embed16(int on) { g_embed16 = on ? 1 : 0; g_alternate = (on == 2) ? 0 : 1; }
void set_gemm_tc16_mode(int on) { g_tc16 = on ? 1 : 0; set_gemm_tc16_pair(on == 2); }

static bool tc16_eligible(const kmpc_model* m) {
  if (!g_tc16 || !g_fold) return false;
  if (m->kind != KMPC_MODEL_GENERIC || m->norm_fn != KMPC_NORM_ID || m->n_dec != 1 || m->n_enc < 1) return false;
  for (int i = 1; i <= m->n_enc; ++i) if (m->enc_dims[i] % 8) return false;
  return true;
}

// fp16-pair copies of the encoder weights and of the folded read-out
static int ensure_tc16(kmpc_handle* h, kmpc_model* m, int H, cudaStream_t st) {
  cudaError_t e;
  if (!m->ovf_flag) {
    if ((e = cudaMalloc(&m->ovf_flag, sizeof(int))) != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(ovf)");
    if ((e = cudaMallocHost(&m->ovf_host, sizeof(int))) != cudaSuccess) return kmpc_fail_cuda(e, "cudaMallocHost(ovf)");
  }
  if (m->w16_hi.empty()) {
    for (int i = 0; i < m->n_enc; ++i) {
      const int out_f = m->enc_dims[i + 1];
      const int in_f = (i == 0) ? m->d * m->ld16 : m->enc_dims[i];
      __half *hi = nullptr, *lo = nullptr;
      if ((e = cudaMalloc(&hi, (size_t)out_f * in_f * sizeof(__half))) != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(w16)");
      if ((e = cudaMalloc(&lo, (size_t)out_f * in_f * sizeof(__half))) != cudaSuccess) { cudaFree(hi); return kmpc_fail_cuda(e, "cudaMalloc(w16)"); }
      m->w16_hi.push_back(hi); m->w16_lo.push_back(lo);
      const float* src = m->enc_w[i];
      float* tmp = nullptr;
      if (i == 0) {          // window layout with row stride ld16
        if ((e = cudaMalloc(&tmp, (size_t)out_f * in_f * sizeof(float))) != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(w16 tmp)");
        window_permute_kernel<<<h->sm_count * 4, 256, 0, st>>>(m->enc_w[0], out_f, m->N, m->d, m->ld16, tmp);
        src = tmp;
      }
      int rc = launch_split16(src, out_f, in_f, in_f, hi, lo, in_f, nullptr, st);
      cudaStreamSynchronize(st);
      if (tmp) cudaFree(tmp);
      if (rc) return kmpc_fail_cuda((cudaError_t)rc, "split16(weights)");
      h->launches += 2;
    }
  }
  if (!m->w16e_hi) {
    const int out_f = m->enc_dims[1], Kp = ((m->d * m->N + 63) / 64) * 64;
    if ((e = cudaMalloc(&m->w16e_hi, (size_t)out_f * Kp * sizeof(__half))) != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(w16e)");
    if ((e = cudaMalloc(&m->w16e_lo, (size_t)out_f * Kp * sizeof(__half))) != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(w16e)");
    int rc = launch_split16(m->enc_w[0], out_f, m->d * m->N, m->d * m->N, m->w16e_hi, m->w16e_lo, Kp, nullptr, st);
    if (rc) return kmpc_fail_cuda((cudaError_t)rc, "split16(first layer)");
    h->launches++;
  }
  if (!m->fold16_hi || m->fold16_H != m->fold_H || m->fold_H < H) {
    if (m->fold16_hi) { cudaFree(m->fold16_hi); cudaFree(m->fold16_lo); m->fold16_hi = m->fold16_lo = nullptr; }
    const size_t rows_pad = (size_t)((m->fold_H * m->N + 127) / 128) * 128;
    if ((e = cudaMalloc(&m->fold16_hi, rows_pad * m->Z * sizeof(__half))) != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(fold16)");
    if ((e = cudaMalloc(&m->fold16_lo, rows_pad * m->Z * sizeof(__half))) != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(fold16)");
    int rc = launch_split16(m->fold_w, (long long)rows_pad, m->Z, m->Z, m->fold16_hi, m->fold16_lo, m->Z, nullptr, st);
    if (rc) return kmpc_fail_cuda((cudaError_t)rc, "split16(fold)");
    m->fold16_H = m->fold_H;
    h->launches++;
  }
  return 0;
}

// GenericKM forecast with fp16-pair operands: encoder MLP + folded read-out, all rows of all paths.  Asynchronous:
// returns 0 once everything is queued (m->ovf_flag is raised on the device if a value left the fp16 range; the caller
// queues the 3xTF32 chain behind it, gated on that flag), 1 when a launch turned out not to be eligible, < 0 on error.
static int run_chain16(kmpc_handle* h, kmpc_model* m, const float* z, int B, int T, int row0, int t0, int t1, int H,
                       const float* std32, const float* mean32, int stat_rows_per_group, float* out, cudaStream_t st) {
  int rc;
  if ((rc = ensure_fold(h, m, H, st))) return rc;
  if ((rc = ensure_tc16(h, m, H, st))) return rc;
  const int Z = m->Z, N = m->N, ld16 = m->ld16, rpp = t1 - t0, M = B * rpp;
  cudaError_t e;
  const int Kp = ((m->d * N + 63) / 64) * 64;
  const bool embed = g_embed16 != 0;
  // (virtual embedding only) the standardised series as an fp16 pair, row stride ld16
  const size_t zn = (size_t)B * T * ld16;
  if (!embed && m->z16_cap < zn) {
    if (m->z16_hi) { cudaFree(m->z16_hi); cudaFree(m->z16_lo); m->z16_hi = m->z16_lo = nullptr; m->z16_cap = 0; }
    if ((e = cudaMalloc(&m->z16_hi, zn * sizeof(__half))) != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(z16)");
    if ((e = cudaMalloc(&m->z16_lo, zn * sizeof(__half))) != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(z16)");
    m->z16_cap = zn;
  }
  cudaMemsetAsync(m->ovf_flag, 0, sizeof(int), st);
  if (!embed) {
    if ((rc = launch_split16(z, (long long)B * T, N, m->ld, m->z16_hi, m->z16_lo, ld16, m->ovf_flag, st)))
      return kmpc_fail_cuda((cudaError_t)rc, "split16(series)");
    h->launches++;
  }
  int maxw = Z;
  for (int v : m->enc_dims) if (v > maxw) maxw = v;
  long long ch = g_chunk_rows;
  if (rpp < M) { ch = (ch / rpp) * rpp; if (ch < rpp) ch = rpp; }
  if (ch > M) ch = M;
  const int CH = (int)ch;
  const size_t act_halves = (size_t)CH * maxw * 4;
  const size_t need = (act_halves + (embed ? (size_t)CH * Kp * 2 : 0)) * sizeof(__half) + 256;
  if (h->scratch_bytes < need) {
    if (h->scratch) cudaFree(h->scratch);
    h->scratch = nullptr; h->scratch_bytes = 0;
    if ((e = cudaMalloc(&h->scratch, need)) != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(forecast scratch)");
    h->scratch_bytes = need;
  }
  __half* act_hi[2]; __half* act_lo[2];
  for (int i = 0; i < 2; ++i) {
    act_hi[i] = (__half*)h->scratch + (size_t)(2 * i) * CH * maxw;
    act_lo[i] = (__half*)h->scratch + (size_t)(2 * i + 1) * CH * maxw;
  }
  __half* emb_hi = (__half*)h->scratch + act_halves;
  __half* emb_lo = emb_hi + (size_t)CH * Kp;
  for (int r0 = 0; r0 < M; r0 += CH) {
    const int rows = (M - r0 < CH) ? (M - r0) : CH;
    int cur = 0;
    if (embed) {
      const int oct = Kp / 8, rpi = oct < 256 ? 256 / oct : 1;
      int blocks = (rows + rpi - 1) / rpi;
      if (blocks > h->sm_count * 16) blocks = h->sm_count * 16;
      embed16_kernel<<<blocks, 256, Kp * sizeof(int), st>>>(z, m->ld, T, rpp, row0 + t0, N, m->d, Kp, r0, rows, emb_hi, emb_lo,
                                                          m->ovf_flag);
      if ((e = cudaGetLastError()) != cudaSuccess) return kmpc_fail_cuda(e, "embed16_kernel");
      h->launches++;
    }
    for (int li = 0; li < m->n_enc; ++li) {
      Gemm16Args g;
      memset(&g, 0, sizeof(g));
      const bool last = (li == m->n_enc - 1);
      if (li == 0 && embed) {
        g.A_hi = emb_hi; g.A_lo = emb_lo; g.a_rows_per_group = rows; g.lda = Kp; g.K = Kp;
      } else if (li == 0) {
        g.A_hi = m->z16_hi + (size_t)(row0 + t0) * ld16; g.A_lo = m->z16_lo + (size_t)(row0 + t0) * ld16;
        g.a_group_stride = (long long)T * ld16; g.a_rows_per_group = rpp; g.lda = ld16; g.row0 = r0; g.K = m->d * ld16;
      } else {
        g.A_hi = act_hi[cur ^ 1]; g.A_lo = act_lo[cur ^ 1]; g.a_rows_per_group = rows; g.lda = m->enc_dims[li]; g.K = m->enc_dims[li];
      }
      g.W_hi = m->w16_hi[li]; g.W_lo = m->w16_lo[li]; g.ldw = g.K;
      if (li == 0 && embed) { g.W_hi = m->w16e_hi; g.W_lo = m->w16e_lo; }
      g.M = rows; g.Nout = m->enc_dims[li + 1]; g.n_store = g.Nout;
      g.bias = m->enc_b[li];
      g.act = last ? (m->enc_last_relu ? EPI_RELU : EPI_NONE) : act_to_epi(m->enc_act);
      g.C16_hi = act_hi[cur]; g.C16_lo = act_lo[cur]; g.ldc16 = g.Nout;
      g.overflow_flag = m->ovf_flag;
      // A layer reads what the previous kernel of the pass has just written (134 MB in, 134 MB out, 126 MB of L2): walking
      // the tiles in the same order as the writer reads the rows that left L2 first.  Every kernel of the pass therefore
      // runs opposite to its predecessor (the gather writes forward).
      g.reverse = (g_alternate && embed) ? ((li & 1) == 0) : 0;
      rc = launch_gemm_tc16(g, st);
      h->launches++;
      if (rc == -100) return 1;                         // not eligible after all: let the TF32 chain take over
      if (rc) return kmpc_fail_cuda((cudaError_t)rc, "gemm_tc16");
      cur ^= 1;
    }
    Gemm16Args d;
    memset(&d, 0, sizeof(d));
    d.A_hi = act_hi[cur ^ 1]; d.A_lo = act_lo[cur ^ 1]; d.a_rows_per_group = rows; d.lda = Z; d.K = Z;
    d.W_hi = m->fold16_hi; d.W_lo = m->fold16_lo; d.ldw = Z;
    d.M = rows; d.Nout = H * N; d.n_store = H * N; d.bias = m->fold_b; d.act = EPI_NONE;
    d.C = out + (size_t)r0 * H * N; d.ldc = (long long)H * N;
    d.std32 = std32; d.mean32 = mean32; d.stat_rows_per_group = stat_rows_per_group; d.stat_ld = N; d.stat_row0 = r0; d.stat_mod = N;
    d.overflow_flag = m->ovf_flag;
    d.reverse = (g_alternate && embed) ? ((m->n_enc & 1) == 0) : 0;
    rc = launch_gemm_tc16(d, st);
    h->launches++;
    if (rc == -100) return 1;
    if (rc) return kmpc_fail_cuda((cudaError_t)rc, "gemm_tc16(read-out)");
  }
  return 0;
}
}  // namespace kmpc

static thread_local char f_err[256];
static int ffail(int code, const char* msg) { snprintf(f_err, sizeof(f_err), "%s", msg); return code; }
#define FCK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return kmpc_fail_cuda(e_, #call); } while (0)

static int dev_copy(kmpc_model* m, const float* src, size_t n, float** out) {
  float* p = nullptr;
  cudaError_t e = cudaMalloc(&p, n * sizeof(float));
  if (e != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(model)");
  m->owned.push_back(p);
  if (src) { e = cudaMemcpy(p, src, n * sizeof(float), cudaMemcpyDeviceToDevice); if (e != cudaSuccess) return kmpc_fail_cuda(e, "copy weights"); }
  *out = p;
  return 0;
}

static int make_lo(kmpc_model* m, const float* src, size_t n, float** out) {
  int rc = dev_copy(m, nullptr, n, out);
  if (rc) return rc;
  rc = kmpc::launch_split_lo(src, *out, (long long)n, 0);
  if (rc) return kmpc_fail_cuda((cudaError_t)rc, "split_lo");
  return 0;
}

extern "C" {

int kmpc_model_free(kmpc_model* m) {
  if (!m) return KMPC_OK;
  kmpc_device_guard dev_guard_(m->h->device);
  for (void* p : m->owned) cudaFree(p);
  if (m->z_lo) cudaFree(m->z_lo);
  if (m->fold_w) { cudaFree(m->fold_w); cudaFree(m->fold_w_lo); }
  if (m->fold_b) cudaFree(m->fold_b);
  for (__half* q : m->w16_hi) cudaFree(q);
  for (__half* q : m->w16_lo) cudaFree(q);
  if (m->w16e_hi) { cudaFree(m->w16e_hi); cudaFree(m->w16e_lo); }
  if (m->fold16_hi) { cudaFree(m->fold16_hi); cudaFree(m->fold16_lo); }
  if (m->z16_hi) { cudaFree(m->z16_hi); cudaFree(m->z16_lo); }
  if (m->ovf_flag) cudaFree(m->ovf_flag);
  if (m->ovf_host) cudaFreeHost(m->ovf_host);
  delete m;
  return KMPC_OK;
}

int kmpc_model_load(kmpc_handle* h, const kmpc_model_desc* D, kmpc_model** out) {
  if (!h || !D || !out) return kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_model_load: NULL argument");
  if (D->obs != D->n_assets * D->delay || D->latent <= 0 || D->n_assets <= 0)
    return kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_model_load: obs must equal n_assets*delay");
  kmpc_device_guard dev_guard_(h->device);
  FCK(dev_guard_.err);
  kmpc_model* m = new kmpc_model();
  m->h = h; m->kind = D->kind; m->obs = D->obs; m->N = D->n_assets; m->d = D->delay; m->Z = D->latent;
  m->ld = ((D->n_assets + 3) / 4) * 4; m->norm_fn = D->norm_fn;
  m->n_enc = 0; m->enc_act = D->enc_act; m->enc_last_relu = D->enc_last_relu; m->enc_w0_win = nullptr;
  m->n_dec = 0; m->dec_act = D->dec_act; m->kmatT = nullptr;
  m->lista_linear = D->lista_linear_encoder; m->lista_loops = D->lista_loops; m->lista_thr = D->lista_threshold;
  m->lista_ST = nullptr; m->lista_wdT = nullptr; m->lista_ST_lo = nullptr; m->lista_wdT_lo = nullptr; m->kmatT_lo = nullptr; m->enc_w0_win_lo = nullptr;
  m->z_lo = nullptr; m->z_lo_cap = 0;
  m->fold_H = 0; m->fold_w = nullptr; m->fold_w_lo = nullptr; m->fold_b = nullptr;
  m->ld16 = ((D->n_assets + 7) / 8) * 8; m->fold16_H = 0; m->fold16_hi = m->fold16_lo = nullptr;
  m->z16_hi = m->z16_lo = nullptr; m->z16_cap = 0; m->ovf_flag = nullptr; m->ovf_host = nullptr;
  int rc = 0;
  auto bail = [&](int code) { kmpc_model_free(m); return code; };
  const bool mlp_enc = (D->kind == KMPC_MODEL_GENERIC) || !D->lista_linear_encoder;
  const float* first_w = nullptr; int first_out = 0;
  if (mlp_enc) {
    if (D->n_enc <= 0 || !D->enc_dims_host || !D->enc_w_host) return bail(kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_model_load: encoder layers missing"));
    m->n_enc = D->n_enc;
    m->enc_dims.assign(D->enc_dims_host, D->enc_dims_host + D->n_enc + 1);
    if (m->enc_dims[0] != D->obs || m->enc_dims[D->n_enc] != D->latent) return bail(kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_model_load: encoder dims must run obs -> latent"));
    for (int i = 0; i < D->n_enc; ++i) {
      float *w = nullptr, *b = nullptr;
      if ((rc = dev_copy(m, D->enc_w_host[i], (size_t)m->enc_dims[i] * m->enc_dims[i + 1], &w))) return bail(rc);
      if (D->enc_b_host && D->enc_b_host[i]) { if ((rc = dev_copy(m, D->enc_b_host[i], m->enc_dims[i + 1], &b))) return bail(rc); }
      float* wl = nullptr;
      if ((rc = make_lo(m, w, (size_t)m->enc_dims[i] * m->enc_dims[i + 1], &wl))) return bail(rc);
      m->enc_w.push_back(w); m->enc_b.push_back(b); m->enc_w_lo.push_back(wl);
    }
    first_w = m->enc_w[0]; first_out = m->enc_dims[1];
  } else {
    if (!D->lista_We) return bail(kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_model_load: lista_We missing"));
    float* w = nullptr;
    if ((rc = dev_copy(m, D->lista_We, (size_t)D->latent * D->obs, &w))) return bail(rc);
    float* wl = nullptr;
    if ((rc = make_lo(m, w, (size_t)D->latent * D->obs, &wl))) return bail(rc);
    m->enc_w.push_back(w); m->enc_b.push_back(nullptr); m->enc_w_lo.push_back(wl);
    m->enc_dims = {D->obs, D->latent};
    first_w = w; first_out = D->latent;
  }
  {  // first layer re-laid for the window read
    if ((rc = dev_copy(m, nullptr, (size_t)first_out * m->d * m->ld, &m->enc_w0_win))) return bail(rc);
    kmpc::window_permute_kernel<<<h->sm_count * 4, 256>>>(first_w, first_out, m->N, m->d, m->ld, m->enc_w0_win);
    if ((rc = make_lo(m, m->enc_w0_win, (size_t)first_out * m->d * m->ld, &m->enc_w0_win_lo))) return bail(rc);
  }
  if (!D->kmat) return bail(kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_model_load: kmat missing"));
  if ((rc = dev_copy(m, nullptr, (size_t)m->Z * m->Z, &m->kmatT))) return bail(rc);
  {
    dim3 grid((m->Z + 31) / 32, (m->Z + 31) / 32), blk(32, 8);
    kmpc::transpose_kernel<<<grid, blk>>>(D->kmat, m->Z, m->Z, m->kmatT);
    if ((rc = make_lo(m, m->kmatT, (size_t)m->Z * m->Z, &m->kmatT_lo))) return bail(rc);
  }
  if (D->kind == KMPC_MODEL_GENERIC) {
    if (D->n_dec <= 0 || !D->dec_dims_host || !D->dec_w_host) return bail(kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_model_load: decoder layers missing"));
    m->n_dec = D->n_dec;
    m->dec_dims.assign(D->dec_dims_host, D->dec_dims_host + D->n_dec + 1);
    if (m->dec_dims[0] != D->latent || m->dec_dims[D->n_dec] != D->obs) return bail(kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_model_load: decoder dims must run latent -> obs"));
    for (int i = 0; i < D->n_dec; ++i) {
      float *w = nullptr, *b = nullptr;
      if ((rc = dev_copy(m, D->dec_w_host[i], (size_t)m->dec_dims[i] * m->dec_dims[i + 1], &w))) return bail(rc);
      if (D->dec_b_host && D->dec_b_host[i]) { if ((rc = dev_copy(m, D->dec_b_host[i], m->dec_dims[i + 1], &b))) return bail(rc); }
      float* wl = nullptr;
      if ((rc = make_lo(m, w, (size_t)m->dec_dims[i] * m->dec_dims[i + 1], &wl))) return bail(rc);
      m->dec_w.push_back(w); m->dec_b.push_back(b); m->dec_w_lo.push_back(wl);
    }
  } else {
    if (!D->lista_S || !D->lista_dict) return bail(kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_model_load: lista_S / lista_dict missing"));
    if ((rc = dev_copy(m, nullptr, (size_t)m->Z * m->Z, &m->lista_ST))) return bail(rc);
    dim3 grid((m->Z + 31) / 32, (m->Z + 31) / 32), blk(32, 8);
    kmpc::transpose_kernel<<<grid, blk>>>(D->lista_S, m->Z, m->Z, m->lista_ST);
    if ((rc = make_lo(m, m->lista_ST, (size_t)m->Z * m->Z, &m->lista_ST_lo))) return bail(rc);
    if ((rc = dev_copy(m, nullptr, (size_t)m->obs * m->Z, &m->lista_wdT))) return bail(rc);
    kmpc::dict_normalize_transpose_kernel<<<(m->Z + 7) / 8, 256>>>(D->lista_dict, m->Z, m->obs, m->lista_wdT);
    if ((rc = make_lo(m, m->lista_wdT, (size_t)m->obs * m->Z, &m->lista_wdT_lo))) return bail(rc);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return bail(kmpc_fail_cuda(e, "kmpc_model_load kernels"));
  h->launches += 3;
  *out = m;
  return KMPC_OK;
}

static int stats_to_f32(kmpc_handle* h, const double* mean, const double* std, int n, float** std32, float** mean32, cudaStream_t st) {
  // small per-call conversion buffer owned by the handle (hence on the handle's device; re-allocated if it grows)
  if (h->stats32_cap < 2 * n) {
    if (h->stats32) cudaFree(h->stats32);
    cudaError_t e = cudaMalloc(&h->stats32, (size_t)2 * n * sizeof(float));
    if (e != cudaSuccess) { h->stats32 = nullptr; h->stats32_cap = 0; return kmpc_fail_cuda(e, "cudaMalloc(stats)"); }
    h->stats32_cap = 2 * n;
  }
  float* buf = h->stats32;
  int blocks = (n + 255) / 256;
  kmpc::f64_to_f32_kernel<<<blocks, 256, 0, st>>>(std, buf, n);
  kmpc::f64_to_f32_kernel<<<blocks, 256, 0, st>>>(mean, buf + n, n);
  h->launches += 2;
  *std32 = buf; *mean32 = buf + n;
  return 0;
}

int kmpc_forecast(kmpc_handle* h, const kmpc_model* m, const float* z, int ld_z, const double* mean, const double* std,
                  int stats_per_path, int B, int T, int row0, int t0, int t1, int H, float* yhat, void* stream) {
  if (!h || !m || !z || !mean || !std || !yhat) return kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_forecast: NULL argument");
  if (ld_z != m->ld) return kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_forecast: ld_z must be n_assets rounded up to a multiple of 4");
  if (B <= 0 || H <= 0 || t1 <= t0 || row0 < 0 || t0 < 0 || row0 + t1 + m->d - 1 > T)
    return kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_forecast: bad row range");
  kmpc_device_guard dev_guard_(h->device);
  FCK(dev_guard_.err);
  cudaStream_t st = (cudaStream_t)stream;
  float *std32, *mean32;
  int rc = stats_to_f32(h, mean, std, (stats_per_path ? B : 1) * m->N, &std32, &mean32, st);
  if (rc) return rc;
  const int rpp = t1 - t0;
  kmpc_model* mm = const_cast<kmpc_model*>(m);
  const int* gate = nullptr;
  if (kmpc::tc16_eligible(m) && rpp >= 1 && B * rpp >= 128) {
    // fp16-pair tensor-core chain (gemm_tc16.cu).  Its epilogues raise mm->ovf_flag on the device when a value leaves
    // the fp16 range (|x| > 65504); the 3xTF32 chain is queued behind it with that flag as its gate and overwrites the
    // forecasts in that case only: no host synchronisation, the call stays asynchronous on `stream`.
    rc = kmpc::run_chain16(h, mm, z, B, T, row0, t0, t1, H, std32, mean32, stats_per_path ? rpp : 0, yhat, st);
    if (rc < 0) return rc;
    if (rc == 0) gate = mm->ovf_flag;
  }
  // residual twin of the series for the 3xTF32 operand split of the first layer
  const size_t zn = (size_t)B * T * ld_z;
  if (mm->z_lo_cap < zn) {
    if (mm->z_lo) cudaFree(mm->z_lo);
    mm->z_lo = nullptr; mm->z_lo_cap = 0;
    cudaError_t e = cudaMalloc(&mm->z_lo, zn * sizeof(float));
    if (e != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(z_lo)");
    mm->z_lo_cap = zn;
  }
  rc = kmpc::launch_split_lo(z, mm->z_lo, (long long)zn, st, gate);
  h->launches++;
  if (rc) return kmpc_fail_cuda((cudaError_t)rc, "split_lo(series)");
  kmpc::AView av;
  av.A = z + (size_t)(row0 + t0) * ld_z; av.A_lo = mm->z_lo + (size_t)(row0 + t0) * ld_z;
  av.group_stride = (long long)T * ld_z; av.rows_per_group = rpp; av.lda = ld_z;
  av.K = m->d * ld_z; av.window = true;
  return kmpc::run_chain(h, m, av, B * rpp, H, 1, m->N, std32, mean32, stats_per_path ? rpp : 0, yhat, st, gate);
}

int kmpc_encode(kmpc_handle* h, const kmpc_model* m, const float* obs, int M, float* latent, void* stream) {
  if (!h || !m || !obs || !latent || M <= 0) return kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_encode: bad argument");
  kmpc_device_guard dev_guard_(h->device);
  FCK(dev_guard_.err);
  kmpc::AView av; av.A = obs; av.A_lo = nullptr; av.group_stride = 0; av.rows_per_group = M; av.lda = m->obs; av.K = m->obs; av.window = false;
  return kmpc::run_chain(h, m, av, M, 0, 0, 0, nullptr, nullptr, 0, latent, (cudaStream_t)stream);
}

int kmpc_rollout(kmpc_handle* h, const kmpc_model* m, const float* obs, int M, int H, int obs_cols, float* pred, void* stream) {
  if (!h || !m || !obs || !pred || M <= 0 || H <= 0 || obs_cols <= 0 || obs_cols > m->obs)
    return kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_rollout: bad argument");
  kmpc_device_guard dev_guard_(h->device);
  FCK(dev_guard_.err);
  kmpc::AView av; av.A = obs; av.A_lo = nullptr; av.group_stride = 0; av.rows_per_group = M; av.lda = m->obs; av.K = m->obs; av.window = false;
  return kmpc::run_chain(h, m, av, M, H, 2, obs_cols, nullptr, nullptr, 0, pred, (cudaStream_t)stream);
}

// KoopmanMachine.step_latent (model.py:311-321, 787-797): out = norm(z @ kmat)
int kmpc_step_latent(kmpc_handle* h, const kmpc_model* m, const float* z, int M, float* out, void* stream) {
  if (!h || !m || !z || !out || M <= 0) return kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_step_latent: bad argument");
  kmpc_device_guard dev_guard_(h->device);
  FCK(dev_guard_.err);
  cudaStream_t st = (cudaStream_t)stream;
  kmpc::GemmArgs g = kmpc::base_args();
  g.A = z; g.a_rows_per_group = M; g.lda = m->Z; g.K = m->Z; g.W = m->kmatT; g.ldw = m->Z;
  g.M = M; g.Nout = m->Z; g.n_store = m->Z; g.C = out; g.ldc = m->Z;
  int rc = kmpc::launch_gemm(g, st, &h->launches);
  if (rc) return rc;
  if (m->kind == KMPC_MODEL_GENERIC && m->norm_fn == KMPC_NORM_BALL) {
    kmpc::row_normalize_kernel<<<(M + 7) / 8, 256, 0, st>>>(out, nullptr, M, m->Z); h->launches++;
  }
  return KMPC_OK;
}

// KoopmanMachine.decode (model.py:768-777, 839-850): out [M, obs]
int kmpc_decode(kmpc_handle* h, const kmpc_model* m, const float* z, int M, float* out, void* stream) {
  if (!h || !m || !z || !out || M <= 0) return kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_decode: bad argument");
  kmpc_device_guard dev_guard_(h->device);
  FCK(dev_guard_.err);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (m->kind == KMPC_MODEL_LISTA) {
    kmpc::GemmArgs d = kmpc::base_args();
    d.A = z; d.a_rows_per_group = M; d.lda = m->Z; d.K = m->Z; d.W = m->lista_wdT; d.ldw = m->Z;
    d.M = M; d.Nout = m->obs; d.n_store = m->obs; d.C = out; d.ldc = m->obs;
    return kmpc::launch_gemm(d, st, &h->launches);
  }
  int maxw = 0;
  for (int v : m->dec_dims) if (v > maxw) maxw = v;
  const size_t need = (size_t)2 * M * maxw * sizeof(float);
  if (m->n_dec > 1 && h->scratch_bytes < need) {
    if (h->scratch) cudaFree(h->scratch);
    h->scratch = nullptr; h->scratch_bytes = 0;
    cudaError_t e = cudaMalloc(&h->scratch, need);
    if (e != cudaSuccess) return kmpc_fail_cuda(e, "cudaMalloc(decode scratch)");
    h->scratch_bytes = need;
  }
  const float* xx = z; int xl = m->Z;
  for (int li = 0; li < m->n_dec; ++li) {
    kmpc::GemmArgs d = kmpc::base_args();
    const bool last = (li == m->n_dec - 1);
    d.A = xx; d.a_rows_per_group = M; d.lda = xl; d.K = m->dec_dims[li]; d.W = m->dec_w[li]; d.ldw = m->dec_dims[li];
    d.M = M; d.bias = m->dec_b[li]; d.Nout = m->dec_dims[li + 1]; d.n_store = d.Nout;
    d.act = last ? kmpc::EPI_NONE : kmpc::act_to_epi(m->dec_act);
    d.C = last ? out : (float*)h->scratch + (size_t)(li & 1) * M * maxw; d.ldc = d.Nout;
    if ((rc = kmpc::launch_gemm(d, st, &h->launches))) return rc;
    xx = d.C; xl = d.Nout;
  }
  return KMPC_OK;
}

// ---- diagnostics -------------------------------------------------------------------------------------------
// 1 = tcgen05 3xTF32 GEMM where eligible (default), 0 = fp32 SIMT GEMM everywhere.  Process-wide.
int kmpc_set_gemm_mode(int use_tensor_cores) { kmpc::set_gemm_tc_mode(use_tensor_cores); return KMPC_OK; }
int kmpc_set_forecast_fold(int on) { kmpc::set_forecast_fold(on); return KMPC_OK; }
int kmpc_set_gemm_fp16_pairs(int on) { kmpc::set_gemm_tc16_mode(on); return KMPC_OK; }
int kmpc_set_forecast_chunk_rows(int rows) { kmpc::set_forecast_chunk_rows(rows); return KMPC_OK; }
int kmpc_set_forecast_embedding(int materialise) { kmpc::set_forecast_embed16(materialise); return KMPC_OK; }

// C[M,Nout] = A[M,K] . W[Nout,K]^T through one chosen kernel: mode 0 = SIMT fp32, 1 = tcgen05 3xTF32 (returns
// KMPC_E_UNSUPPORTED if the shape is not eligible).  Allocates the residual twins internally; synchronous.
int kmpc_debug_gemm(kmpc_handle* h, const float* A, const float* W, int M, int Nout, int K, float* C, int mode) {
  if (!h || !A || !W || !C || M <= 0 || Nout <= 0 || K <= 0) return kmpc_fail_cuda(cudaErrorInvalidValue, "kmpc_debug_gemm: bad argument");
  kmpc_device_guard dev_guard_(h->device);
  FCK(dev_guard_.err);
  if (mode == 2) {           // fp16-pair tensor-core kernel
    __half *Ah = nullptr, *Al = nullptr, *Wh = nullptr, *Wl = nullptr;
    FCK(cudaMalloc(&Ah, (size_t)M * K * sizeof(__half))); FCK(cudaMalloc(&Al, (size_t)M * K * sizeof(__half)));
    FCK(cudaMalloc(&Wh, (size_t)Nout * K * sizeof(__half))); FCK(cudaMalloc(&Wl, (size_t)Nout * K * sizeof(__half)));
    kmpc::launch_split16(A, M, K, K, Ah, Al, K, nullptr, 0);
    kmpc::launch_split16(W, Nout, K, K, Wh, Wl, K, nullptr, 0);
    kmpc::Gemm16Args g;
    memset(&g, 0, sizeof(g));
    g.A_hi = Ah; g.A_lo = Al; g.a_rows_per_group = M; g.lda = K; g.K = K; g.W_hi = Wh; g.W_lo = Wl; g.ldw = K;
    g.M = M; g.Nout = Nout; g.n_store = Nout; g.C = C; g.ldc = Nout; g.act = kmpc::EPI_NONE;
    int rc = kmpc::launch_gemm_tc16(g, 0);
    h->launches += 3;
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(Ah); cudaFree(Al); cudaFree(Wh); cudaFree(Wl);
    if (rc == -100) return kmpc_fail_cuda(cudaErrorNotSupported, "kmpc_debug_gemm: shape not eligible for the fp16-pair tcgen05 kernel");
    if (rc) return kmpc_fail_cuda((cudaError_t)rc, "kmpc_debug_gemm launch");
    if (e != cudaSuccess) return kmpc_fail_cuda(e, "kmpc_debug_gemm kernel");
    return KMPC_OK;
  }
  float *Alo = nullptr, *Wlo = nullptr;
  FCK(cudaMalloc(&Alo, (size_t)M * K * sizeof(float)));
  cudaError_t e = cudaMalloc(&Wlo, (size_t)Nout * K * sizeof(float));
  if (e != cudaSuccess) { cudaFree(Alo); return kmpc_fail_cuda(e, "cudaMalloc"); }
  kmpc::launch_split_lo(A, Alo, (long long)M * K, 0);
  kmpc::launch_split_lo(W, Wlo, (long long)Nout * K, 0);
  kmpc::GemmArgs g = kmpc::base_args();
  g.A = A; g.A_lo = Alo; g.a_rows_per_group = M; g.lda = K; g.K = K; g.W = W; g.W_lo = Wlo; g.ldw = K;
  g.M = M; g.Nout = Nout; g.n_store = Nout; g.C = C; g.ldc = Nout;
  int rc = mode ? kmpc::launch_gemm_tc(g, 0) : kmpc::launch_gemm_simt(g, 0);
  h->launches += 3;
  e = cudaDeviceSynchronize();
  cudaFree(Alo); cudaFree(Wlo);
  if (rc == -100) return kmpc_fail_cuda(cudaErrorNotSupported, "kmpc_debug_gemm: shape not eligible for the tcgen05 kernel");
  if (rc) return kmpc_fail_cuda((cudaError_t)rc, "kmpc_debug_gemm launch");
  if (e != cudaSuccess) return kmpc_fail_cuda(e, "kmpc_debug_gemm kernel");
  return KMPC_OK;
}

}  // extern "C"
