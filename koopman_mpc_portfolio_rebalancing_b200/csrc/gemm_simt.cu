// fp32 SIMT GEMM with the fused epilogues of the forecast path.  128x128x16 tiles, 256 threads, 8x8
// register micro-tiles, register-prefetch double buffering.  This is the shape-agnostic path (ragged
// sizes, tiny decoders, parity hooks); the bulk layers go through the tcgen05 kernel in gemm_tc.cu.
#include "gemm.cuh"

namespace kmpc {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, NT = 256;

__device__ __forceinline__ const float* a_row_ptr(const GemmArgs& g, int m) {
  m += g.row0;
  const int grp = m / g.a_rows_per_group;
  const int r = m - grp * g.a_rows_per_group;
  return g.A + (long long)grp * g.a_group_stride + (long long)r * g.lda;
}

template <bool VEC>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(GemmArgs g) {
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Ws[2][BK][BN + 4];
  const int tid = threadIdx.x;
  if (g.gate && *g.gate == 0) return;           // gated launch (see GemmArgs::gate)
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // loader mapping: each thread loads 2 rows x 4 k of A and of W per k-tile (128 rows x 16 k = 512 float4)
  const int lr = tid >> 2;            // 0..63
  const int lk = (tid & 3) * 4;       // 0,4,8,12
  const float* arow[2];
  const float* wrow[2];
  bool aok[2], wok[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int m = m0 + lr + 64 * i, n = n0 + lr + 64 * i;
    aok[i] = m < g.M; wok[i] = n < g.Nout;
    arow[i] = aok[i] ? a_row_ptr(g, m) : g.A;
    wrow[i] = wok[i] ? g.W + (long long)n * g.ldw : g.W;
  }
  float4 ra[2], rw[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = k0 + lk;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vw = va;
      if (VEC) {
        if (aok[i] && k < g.K) va = *reinterpret_cast<const float4*>(arow[i] + k);
        if (wok[i] && k < g.K) vw = *reinterpret_cast<const float4*>(wrow[i] + k);
      } else {
        if (aok[i]) {
          if (k + 0 < g.K) va.x = arow[i][k + 0];
          if (k + 1 < g.K) va.y = arow[i][k + 1];
          if (k + 2 < g.K) va.z = arow[i][k + 2];
          if (k + 3 < g.K) va.w = arow[i][k + 3];
        }
        if (wok[i]) {
          if (k + 0 < g.K) vw.x = wrow[i][k + 0];
          if (k + 1 < g.K) vw.y = wrow[i][k + 1];
          if (k + 2 < g.K) vw.z = wrow[i][k + 2];
          if (k + 3 < g.K) vw.w = wrow[i][k + 3];
        }
      }
      ra[i] = va; rw[i] = vw;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int r = lr + 64 * i;
      As[buf][lk + 0][r] = ra[i].x; As[buf][lk + 1][r] = ra[i].y; As[buf][lk + 2][r] = ra[i].z; As[buf][lk + 3][r] = ra[i].w;
      Ws[buf][lk + 0][r] = rw[i].x; Ws[buf][lk + 1][r] = rw[i].y; Ws[buf][lk + 2][r] = rw[i].z; Ws[buf][lk + 3][r] = rw[i].w;
    }
  };
  const int tx = tid & 15, ty = tid >> 4;        // 16 x 16 threads, each 8 x 8 outputs (strided by 16 -> conflict-free)
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  const int nk = (g.K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[TM], wv[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4 + i * 16]);
        av[i] = t.x; av[i + 1] = t.y; av[i + 2] = t.z; av[i + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&Ws[buf][kk][tx * 4 + j * 16]);
        wv[j] = t.x; wv[j + 1] = t.y; wv[j + 2] = t.z; wv[j + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  // epilogue: rows ty*4 + (i/4)*64 + i%4, cols tx*4 + (j/4)*64 + j%4
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * 4 + (i >> 2) * 64 + (i & 3);
    if (m >= g.M) continue;
    const int sg = (g.std32 && g.stat_rows_per_group > 0) ? (m + g.stat_row0) / g.stat_rows_per_group : 0;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * 4 + (j >> 2) * 64 + (j & 3);
      if (n >= g.Nout || n >= g.n_store) continue;
      float x = acc[i][j];
      if (g.bias) x += g.bias[n];
      if (g.addend) x += g.addend[(long long)m * g.ld_add + n];
      x = epilogue_apply(x, g.act, g.shrink_thr);
      if (g.std32) {
        const int sn = g.stat_mod ? n % g.stat_mod : n;
        x = __fadd_rn(__fmul_rn(x, g.std32[(long long)sg * g.stat_ld + sn]), g.mean32[(long long)sg * g.stat_ld + sn]);
      }
      g.C[(long long)m * g.ldc + n] = x;
      if (g.C_lo) g.C_lo[(long long)m * g.ldc + n] = tf32_residual(x);
    }
  }
}

int launch_gemm_simt(const GemmArgs& g, cudaStream_t st) {
  if (g.M <= 0 || g.Nout <= 0) return 0;
  dim3 grid((g.Nout + BN - 1) / BN, (g.M + BM - 1) / BM);
  const bool vec = (g.lda % 4 == 0) && (g.ldw % 4 == 0) && (g.K % 4 == 0) && (g.a_group_stride % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(g.A) & 15) == 0) && ((reinterpret_cast<uintptr_t>(g.W) & 15) == 0);
  if (vec) gemm_simt_kernel<true><<<grid, NT, 0, st>>>(g);
  else gemm_simt_kernel<false><<<grid, NT, 0, st>>>(g);
  return (int)cudaGetLastError();
}

}  // namespace kmpc
