// GEMM interface of the forecast path:  C = epilogue(A[M,K] . W[Nout,K]^T)
// (nn.Linear convention: weight [out,in] row-major, x @ W^T + b, model.py:96-117).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "per_device.cuh"

namespace kmpc {

enum : int { EPI_NONE = 0, EPI_RELU = 1, EPI_TANH = 2, EPI_GELU = 3, EPI_SHRINK = 4 };

struct GemmArgs {
  // A row m starts at A + (m / a_rows_per_group) * a_group_stride + (m % a_rows_per_group) * lda  (elements).
  // A plain matrix uses a_rows_per_group = M.  The delay-embedded window view of a standardised series
  // [B,T,ld] uses rows_per_group = rows per path, group_stride = T*ld, lda = ld, K = d*ld: consecutive rows
  // overlap and the embedded matrix is never materialised.
  const float* A;
  const float* A_lo;   // fp32 residual twin of A (x - top19bits(x)); null -> the tcgen05 path is not eligible
  long long a_group_stride;
  int a_rows_per_group;
  int lda;
  int row0;            // logical index of row 0 of this launch (chunked launches over a grouped view)
  const float* W;      // [Nout, K], row stride ldw
  const float* W_lo;   // residual twin of W
  int ldw;
  int M, Nout, K;
  const float* bias;   // [Nout] or null
  const float* addend; // [M, ld_add] added before the activation (LISTA: z @ S + c), or null
  int ld_add;
  int act;             // EPI_*
  float shrink_thr;    // EPI_SHRINK: sign(x) * max(|x| - thr, 0)   (model.py:30-40)
  // de-standardise epilogue (data_finance.py:740-742): out = fl(fl(x * std32[g,col]) + mean32[g,col]),
  // g = m / stat_rows_per_group (0 = shared stats).  Enabled when std32 != null.
  const float* std32;
  const float* mean32;
  int stat_rows_per_group;
  int stat_ld;
  int stat_row0;       // logical index of row 0 of this launch for the statistics lookup
  int stat_mod;        // > 0: statistics column = output column % stat_mod (folded multi-horizon output [H*N])
  float* C;            // output, row stride ldc, only columns < n_store are written
  long long ldc;
  int n_store;
  // optional second output holding the fp32 residual of the TF32 rounding of C (3xTF32 operand split)
  float* C_lo;
  // optional device flag: the kernel does nothing unless *gate != 0 (the 3xTF32 chain queued behind the fp16-pair
  // chain runs only when that chain raised its range flag; no host synchronisation in between)
  const int* gate;
};

// fp16-pair tensor-core path (gemm_tc16.cu): an fp32 value x travels as hi = fp16(x), lo = fp16((x - hi) * 2^11)
struct Gemm16Args {
  const __half* A_hi; const __half* A_lo;   // row addressing as in GemmArgs (elements are halves)
  long long a_group_stride;
  int a_rows_per_group, lda, row0;
  const __half* W_hi; const __half* W_lo;   // [Nout, K], row stride ldw
  int ldw;
  int M, Nout, K;
  const float* bias;
  int act;                                   // EPI_NONE / RELU / TANH / GELU
  __half* C16_hi; __half* C16_lo;            // fp16-pair output (next layer's operand) or null
  long long ldc16;
  float* C;                                  // fp32 output or null; row stride ldc, columns < n_store
  long long ldc;
  int n_store;
  const float* std32; const float* mean32;   // de-standardise epilogue on the fp32 output, as in GemmArgs
  int stat_rows_per_group, stat_ld, stat_row0, stat_mod;
  int* overflow_flag;                        // device int, set to 1 when a value outside the fp16 range is written
  int reverse;                               // 1: walk the output tiles from the last row block to the first (see run_chain16)
};
// returns -100 when the launch is not eligible (shape, alignment)
int launch_gemm_tc16(const Gemm16Args& g, cudaStream_t st);
void set_gemm_tc16_pair(int on);
int launch_split16(const float* x, long long rows, int cols, int ld_in, __half* hi, __half* lo, int ld_out, int* overflow,
                   cudaStream_t st);

// fp32 SIMT path (exact fp32 FMA accumulation; any shape / alignment)
int launch_gemm_simt(const GemmArgs& g, cudaStream_t st);
// tcgen05 3xTF32 path; returns -100 when the launch is not eligible (shape, alignment, missing twins)
int launch_gemm_tc(const GemmArgs& g, cudaStream_t st);
void set_gemm_tc_mode(int on);
int launch_split_lo(const float* x, float* lo, long long n, cudaStream_t st, const int* gate = nullptr);
int launch_gemm(const GemmArgs& g, cudaStream_t st, long long* launches);

// 3xTF32 operand split: kind::tf32 reads the top 19 bits of an fp32 container, x_hi = trunc19(x), and the residual twin
// carries x - x_hi (exact in fp32, <= 13 significant bits).  The MMA would TRUNCATE that residual to TF32 as well — a
// one-sided error of up to 2^-21 |x| that adds up over K (the step-by-step 3xTF32 chain sat at 1.0e-5 of the reference
// on the config-2 golden); rounding the residual to nearest here halves the bound and removes the bias.
__device__ __forceinline__ float tf32_residual(float x) {
  const float r = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(r));
  return __uint_as_float(u);
}

__device__ __forceinline__ float epilogue_apply(float x, int act, float thr) {
  switch (act) {
    case EPI_RELU: return (x < 0.0f) ? 0.0f : x;               // NaN propagates like torch.relu
    case EPI_TANH: return tanhf(x);
    case EPI_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
    case EPI_SHRINK: {
      if (x != x) return x;                                       // sign(nan) * maximum(nan, 0) = nan
      const float m = fmaxf(fabsf(x) - thr, 0.0f);
      return (x > 0.0f) ? m : ((x < 0.0f) ? -m : 0.0f * m);
    }
    default: return x;
  }
}

}  // namespace kmpc
