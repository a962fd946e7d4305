// tcgen05 / TMA GEMM with fp32-class accuracy from FP16 operand pairs, for the MLP encoder and the folded read-out of
// the forecast path:    C[M, Nout] = epilogue( A[M,K] . W[Nout,K]^T )
//
// Why a second tensor-core kernel: the 3xTF32 kernel (gemm_tc.cu) moves 8 bytes per operand element (x and its
// residual twin, fp32 each); at 128x128x32 tiles that is 64 KB per k-block and SM, and the forecast chain pulls
// 8.2 TB/s out of L2 (ncu: lts__t_sectors) with the tensor pipe 34 % busy.  Here an fp32 value x travels as two
// halves:  hi = fp16(x)  and  lo = fp16((x - hi) * 2^11)  — 22 significant bits like the TF32 pair, 4 bytes per
// element, and kind::f16 MMAs run at twice the TF32 rate.  D = A_hi.W_hi (three rotating fp32 TMEM accumulators,
// see gemm_tc.cu on accumulation truncation) + 2^-11 (A_lo.W_hi + A_hi.W_lo) (fourth accumulator).
// Range: |x| must stay below 65504; the epilogue raises a device flag when it writes a half that overflowed and the
// caller re-runs the TF32 chain (forecast.cu).  Values below 6e-5 keep an absolute accuracy of 2^-25.
//
// Same structure as gemm_tc.cu: one CTA per SM, persistent over 128 x 128 output tiles, k-blocks of 64 halves
// (= one 128-byte swizzle row), 3-stage TMA ring of (A_hi, A_lo, W_hi, W_lo) boxes, warp 0 TMA producer, warp 1
// single-thread MMA issuer (12 tcgen05.mma.kind::f16 per stage), warp 2 TMEM allocator, warps 4..11 epilogue: two per
// TMEM lane quarter, half of the columns each.  The four accumulators fill TMEM; the epilogue warps drain them into
// registers, release them, and store while the next tile's MMAs run (epilogue_tile).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "gemm.cuh"

namespace kmpc {

namespace tc16 {

constexpr int BM = 128, BN = 128, BK = 64;          // BK halves = 128 bytes = one swizzle atom row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 2;             // 16 KB (BM == BN)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;         // A_hi, A_lo, W_hi, W_lo
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 8;                       // two warps per TMEM lane quarter, half of the tile's columns each
constexpr int NUM_THREADS = 32 * (EPI_WARP0 + EPI_WARPS);
constexpr int TMEM_COLS = 512;                      // 3 hi accumulators + 1 lo accumulator of 128 fp32 columns
constexpr int NUM_HI = 3;
constexpr uint32_t SPIN_LIMIT = 1u << 28;           // bounded waits: trap instead of hanging the GPU

struct Params {
  int M, Nout, K;
  int rows_per_group, tiles_per_group, n_groups;     // M = n_groups * rows_per_group
  int tiles_n, num_tiles, k_blocks;
  int reverse;                                       // tile t stands for tile num_tiles - 1 - t
  const float* bias;
  int act;
  const float* std32; const float* mean32; int stat_rows_per_group; int stat_ld; int row0; int stat_mod;
  __half* C16_hi; __half* C16_lo; long long ldc16;
  float* C; long long ldc; int n_store;
  int* overflow;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > SPIN_LIMIT) { printf("kmpc gemm_tc16: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// shared-memory matrix descriptor, K-major, SWIZZLE_128B: start>>4 | SBO(1024 B)>>4 at bit 32 | version 1 at bit 46 |
// layout SWIZZLE_128B (=2) at bit 61   (cute/arch/mma_sm100_desc.hpp, UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (UMMA::InstrDescriptor): c = F32 (1 at bit 4), a = b = F16 (0 at bits 7 / 10), K-major both,
// N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int ACT>
__device__ __forceinline__ float act_apply(float x) {
  if (ACT == EPI_RELU) return (x < 0.0f) ? 0.0f : x;                       // NaN propagates like torch.relu
  if (ACT == EPI_TANH) return tanhf(x);
  if (ACT == EPI_GELU) return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
  return x;
}

// Interior block of the fp16-pair output: 16 full columns, bias + activation, branch-free.
template <int ACT>
__device__ __forceinline__ void store_pair_block(const Params& p, const float (&acc)[16], long long m, int n0) {
  float x[16];
  if (p.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);          // warp-uniform address: broadcast loads
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      const float4 b = b4[j4];
      x[j4 * 4] = acc[j4 * 4] + b.x; x[j4 * 4 + 1] = acc[j4 * 4 + 1] + b.y;
      x[j4 * 4 + 2] = acc[j4 * 4 + 2] + b.z; x[j4 * 4 + 3] = acc[j4 * 4 + 3] + b.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = acc[j];
  }
  __align__(16) __half2 hh[8];
  __align__(16) __half2 ll[8];
  float mx = 0.0f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float a = act_apply<ACT>(x[2 * j]), b = act_apply<ACT>(x[2 * j + 1]);
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    hh[j] = h;
    ll[j] = __floats2half2_rn((a - hf.x) * 2048.0f, (b - hf.y) * 2048.0f);
    mx = fmaxf(mx, fmaxf(fabsf(a), fabsf(b)));
  }
  if (!(mx <= 65504.0f) && p.overflow) *p.overflow = 1;                       // also catches NaN
  uint4* hrow = reinterpret_cast<uint4*>(p.C16_hi + m * p.ldc16 + n0);
  uint4* lrow = reinterpret_cast<uint4*>(p.C16_lo + m * p.ldc16 + n0);
  hrow[0] = reinterpret_cast<const uint4*>(hh)[0]; hrow[1] = reinterpret_cast<const uint4*>(hh)[1];
  lrow[0] = reinterpret_cast<const uint4*>(ll)[0]; lrow[1] = reinterpret_cast<const uint4*>(ll)[1];
}

// Epilogue of one 128 x 128 accumulator set for one warp: TMEM lane quarter q (rows q*32 + lane of the CTA's tile),
// column blocks [cb0, cb0 + CB_PER_WARP) of 16.  m = global output row of this lane, tn = column tile.
// Two phases.  DRAIN: the warp's share of the four accumulators is read out of TMEM and combined into 16 x CB_PER_WARP
// fp32 registers per thread, then `release()` hands the accumulators back to the MMA issuer.  STORE: bias /
// activation / fp16 split / global stores run from registers while the next tile's MMAs are already accumulating —
// the four accumulators fill TMEM, so they cannot be double-buffered; the registers of the epilogue warps are the
// second buffer.  (Before: the accumulators were held through the stores: layers 2-3 took 197 instead of 170 us per chunk.)
#ifndef KMPC_TC16_EARLY_RELEASE
#define KMPC_TC16_EARLY_RELEASE 1
#endif
constexpr int CB_PER_WARP = (BN / 16) / (EPI_WARPS / 4);       // column blocks of 16 per warp
__device__ __forceinline__ void drain_block(uint32_t lane_addr, int cb, int nhi, float (&acc)[16]) {
  uint32_t v[16], u0[16], u1[16], u2[16];
  tmem_ld16_nowait(lane_addr + (uint32_t)(NUM_HI * BN + cb * 16), v);    // lo products, scaled by 2^11
  tmem_ld16_nowait(lane_addr + (uint32_t)(0 * BN + cb * 16), u0);
  if (nhi > 1) tmem_ld16_nowait(lane_addr + (uint32_t)(1 * BN + cb * 16), u1);
  if (nhi > 2) tmem_ld16_nowait(lane_addr + (uint32_t)(2 * BN + cb * 16), u2);
  tmem_wait_ld();
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float a = __fadd_rn(__uint_as_float(v[j]) * (1.0f / 2048.0f), __uint_as_float(u0[j]));
    if (nhi > 1) a = __fadd_rn(a, __uint_as_float(u1[j]));
    if (nhi > 2) a = __fadd_rn(a, __uint_as_float(u2[j]));
    acc[j] = a;
  }
}

// Every block that is not an interior fp16-pair block with a cheap activation (fp32 output: read-out, encode / decode
// hooks; tanh / gelu layers; ragged column edges) goes through a per-warp staging tile ([32][17] floats, conflict-free
// both ways) and a ROLLED loop over its rows (one copy of the element code per block).  Two reasons, both measured
// (profiles/README.md):
//  * code size.  With bias / activation switch (inlined tanhf, erff) / de-standardisation unrolled over the 16 elements
//    of a block the kernel was 27 800 SASS instructions (445 KB): every tile's epilogue streamed its code from L2, and
//    the read-out ran at 28-52 us per tile against a 12 us main loop (tensor pipe 10 %);
//  * store pattern.  A thread owns one ROW of the accumulator, so a store from the accumulator registers puts every lane
//    on a different 128-byte line (rows are 1000 bytes apart in the config-2 read-out).  From the tile the block is stored
//    TRANSPOSED: lane = (row parity, column), 16 instructions of two contiguous row segments each.
// The per-column operands (bias, statistics column) are loaded once per block; the statistics row changes at most once
// inside a warp's 32 rows when a group has >= 32 rows (else it is looked up per row).
constexpr int STAGE_LD = 17;
constexpr int STAGE_FLOATS_PER_WARP = 32 * STAGE_LD;
constexpr int BAR_BYTES = 256;                       // mbarriers + TMEM slot, after the operand stages
constexpr int EPI_STAGE_BYTES = EPI_WARPS * 32 * STAGE_LD * 4;
// tanh / gelu / shrink: one out-of-line copy (inlined per element they were most of the kernel's code)
__device__ __noinline__ float act_slow(float x, int act) { return epilogue_apply(x, act, 0.0f); }

__device__ __forceinline__ void store_rows_staged(const Params& p, const float* stage, int n0, long long m0, int rows_ok, int lane) {
  const int col = lane & 15, half = lane >> 4;
  const int n = n0 + col;
  if (n >= p.n_store) return;
  // every field used inside the row loop is copied to a local first: the loop stores through pointers, so the compiler
  // would otherwise re-load the fields from the parameter block in every iteration (measured: six dependent generic
  // loads per element, 23 us of epilogue per tile)
  float* const C = p.C; __half* const C16_hi = p.C16_hi; __half* const C16_lo = p.C16_lo;
  const long long ldc = p.ldc, ldc16 = p.ldc16;
  const int act = p.act;
  const bool has_bias = p.bias != nullptr;
  const float b = has_bias ? p.bias[n] : 0.0f;
  const float* const std32 = p.std32; const float* const mean32 = p.mean32;
  const bool stats = std32 != nullptr;
  const int stat_ld = p.stat_ld;
  const long long row0 = p.row0;
  const int sn = p.stat_mod ? n % p.stat_mod : n;
  const int srpg = p.stat_rows_per_group;
  long long sg0 = 0; int rem0 = 0;
  float s0 = 1.0f, mu0 = 0.0f, s1 = 1.0f, mu1 = 0.0f;
  const bool two = stats && srpg >= 32;               // at most two statistics rows in this warp's 32 rows
  if (stats && srpg > 0) { sg0 = (m0 + row0) / srpg; rem0 = (int)((m0 + row0) - sg0 * srpg); }
  if (stats && (two || srpg <= 0)) {
    s0 = std32[sg0 * stat_ld + sn]; mu0 = mean32[sg0 * stat_ld + sn];
    if (two && rem0 + rows_ok > srpg) { s1 = std32[(sg0 + 1) * stat_ld + sn]; mu1 = mean32[(sg0 + 1) * stat_ld + sn]; }
  }
  bool ovf = false;
#pragma unroll 4
  for (int r = half; r < rows_ok; r += 2) {
    float t = stage[r * STAGE_LD + col];
    if (has_bias) t += b;
    if (act == EPI_RELU) t = (t < 0.0f) ? 0.0f : t;              // NaN propagates like torch.relu
    else if (act != EPI_NONE) t = act_slow(t, act);
    if (stats) {
      float sd = s0, mn = mu0;
      if (two) { if (rem0 + r >= srpg) { sd = s1; mn = mu1; } }
      else if (srpg > 0) {
        const long long sg = (m0 + r + row0) / srpg;
        sd = std32[sg * stat_ld + sn]; mn = mean32[sg * stat_ld + sn];
      }
      t = __fadd_rn(__fmul_rn(t, sd), mn);
    }
    const long long m = m0 + r;
    if (C16_hi) {                       // next layer's operand: fp16 pair, lo scaled by 2^11
      const __half h = __float2half_rn(t);
      C16_hi[m * ldc16 + n] = h;
      C16_lo[m * ldc16 + n] = __float2half_rn((t - __half2float(h)) * 2048.0f);
      ovf = ovf || !(fabsf(t) <= 65504.0f);
    }
    if (C) C[m * ldc + n] = t;
  }
  if (ovf && p.overflow) *p.overflow = 1;
}

// Interior block of an fp32 output (read-out, encode / decode hooks): 16 full columns straight from the accumulator
// registers.  The staging tile above costs shared-memory bandwidth that the tensor core and TMA are already using (the
// kernel's busiest unit, ~85 % in the main loop): measured 15 us of epilogue per tile through the tile against a 12 us
// main loop.  Here every lane stores its own row's 64 bytes with the widest vectors the WARP's rows all allow (rows of
// the config-2 read-out are 1000 bytes apart: 8-byte aligned) — scattered over 32 lines, but fire-and-forget.
__device__ __forceinline__ void store_c_block(const Params& p, const float (&acc)[16], long long m, int n0, bool row_ok,
                                              long long sg) {
  float x[16];
  if (p.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);          // warp-uniform address: broadcast loads
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      const float4 b = b4[j4];
      x[j4 * 4] = acc[j4 * 4] + b.x; x[j4 * 4 + 1] = acc[j4 * 4 + 1] + b.y;
      x[j4 * 4 + 2] = acc[j4 * 4 + 2] + b.z; x[j4 * 4 + 3] = acc[j4 * 4 + 3] + b.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = acc[j];
  }
  if (p.act == EPI_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = (x[j] < 0.0f) ? 0.0f : x[j];
  }
  if (p.std32) {
    const float* sd = p.std32 + sg * p.stat_ld;
    const float* mn = p.mean32 + sg * p.stat_ld;
    const int smod = p.stat_mod ? p.stat_mod : 0x7fffffff;
    int sn = p.stat_mod ? n0 % p.stat_mod : n0;                    // statistics column of output column n
    if (!row_ok) sn = 0;                                           // rows beyond the group read (and discard) row sg's first entries
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      x[j] = __fadd_rn(__fmul_rn(x[j], sd[sn]), mn[sn]);
      if (++sn == smod) sn = 0;
    }
  }
  float* crow = p.C + m * p.ldc + n0;
  const unsigned a16 = __all_sync(0xffffffffu, !row_ok || (((uintptr_t)crow) & 15) == 0);
  const unsigned a8 = __all_sync(0xffffffffu, !row_ok || (((uintptr_t)crow) & 7) == 0);
  if (!row_ok) return;
  if (a16) {
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4)
      *reinterpret_cast<float4*>(crow + j4 * 4) = make_float4(x[j4 * 4], x[j4 * 4 + 1], x[j4 * 4 + 2], x[j4 * 4 + 3]);
  } else if (a8) {
#pragma unroll
    for (int j2 = 0; j2 < 8; ++j2) *reinterpret_cast<float2*>(crow + j2 * 2) = make_float2(x[j2 * 2], x[j2 * 2 + 1]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) crow[j] = x[j];
  }
}

// `release()` is called exactly once, by all lanes of the warp, when the warp no longer needs the accumulators.
// `stage`: this warp's STAGE_FLOATS_PER_WARP floats of shared memory.
template <typename Release>
__device__ __forceinline__ void epilogue_tile(const Params& p, uint32_t tmem_base, int q, int cb0, int tn, long long m,
                                              bool row_ok, int nhi, float* stage, Release release) {
  const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
  const int lane = threadIdx.x & 31;
  const int rows_ok = __popc(__ballot_sync(0xffffffffu, row_ok));  // valid rows are a prefix of the warp's 32
  const long long m0 = m - lane;
  const bool cheap_act = (p.act == EPI_RELU || p.act == EPI_NONE);
  const bool pair_fast = p.C16_hi && !p.C && cheap_act;           // warp-uniform
  const bool c_fast = p.C && !p.C16_hi && cheap_act;
  const int srpg = p.stat_rows_per_group;
  const long long sg = (c_fast && p.std32 && srpg > 0 && row_ok) ? (m + p.row0) / srpg : 0;   // my row's statistics row
  auto store = [&](const float (&a)[16], int cb) {
    const int n0 = tn * BN + cb * 16;
    if (pair_fast && n0 + 16 <= p.n_store) {                     // interior block of an fp16-pair layer
      if (row_ok) {
        if (p.act == EPI_RELU) store_pair_block<EPI_RELU>(p, a, m, n0);
        else store_pair_block<EPI_NONE>(p, a, m, n0);
      }
    } else if (c_fast && n0 + 16 <= p.n_store) {                 // interior block of an fp32 output
      store_c_block(p, a, m, n0, row_ok, sg);
    } else if (n0 < p.n_store) {
#pragma unroll
      for (int j = 0; j < 16; ++j) stage[lane * STAGE_LD + j] = a[j];
      __syncwarp();
      store_rows_staged(p, stage, n0, m0, rows_ok, lane);
      __syncwarp();                                              // the staging tile is reused by the next block
    }
  };
#if KMPC_TC16_EARLY_RELEASE
  float acc[CB_PER_WARP][16];
#pragma unroll
  for (int c = 0; c < CB_PER_WARP; ++c) drain_block(lane_addr, cb0 + c, nhi, acc[c]);
  release();
#pragma unroll
  for (int c = 0; c < CB_PER_WARP; ++c) store(acc[c], cb0 + c);
#else
#pragma unroll 1
  for (int cb = cb0; cb < cb0 + CB_PER_WARP; ++cb) {
    float acc[16];
    drain_block(lane_addr, cb, nhi, acc);
    store(acc, cb);
  }
  release();
#endif
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc16_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapWlo,
                 const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: stages first (1024-aligned), then barriers
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(base + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;     // [1] accumulators ready
  uint64_t* tempty_bar = tfull_bar + 1;         // [1] accumulators drained
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* stage = (float*)(base + STAGES * STAGE_BYTES + BAR_BYTES) + (warp >= EPI_WARP0 ? warp - EPI_WARP0 : 0) * STAGE_FLOATS_PER_WARP;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAlo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapWlo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar[0], 1); mbar_init(&tempty_bar[0], EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int te = p.reverse ? p.num_tiles - 1 - tile : tile;
        const int tn = te % p.tiles_n, tm = te / p.tiles_n;
        const int g = tm / p.tiles_per_group, tb = tm - g * p.tiles_per_group;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = base + stage * STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_3d(&mapA, &full_bar[stage], st, kb * BK, tb * BM, g);
          tma_load_3d(&mapAlo, &full_bar[stage], st + TILE_BYTES, kb * BK, tb * BM, g);
          tma_load_2d(&mapW, &full_bar[stage], st + 2 * TILE_BYTES, kb * BK, tn * BN);
          tma_load_2d(&mapWlo, &full_bar[stage], st + 3 * TILE_BYTES, kb * BK, tn * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[0], acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_lo = tmem_base + NUM_HI * BN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(base + stage * STAGE_BYTES);
          const uint64_t dA = make_desc(sa), dAlo = make_desc(sa + TILE_BYTES);
          const uint64_t dW = make_desc(sa + 2 * TILE_BYTES), dWlo = make_desc(sa + 3 * TILE_BYTES);
          const uint32_t d_hi = tmem_base + (uint32_t)(kb % NUM_HI) * BN;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t adv = (uint64_t)((kk * 32) >> 4);       // 16 halves = 32 bytes inside the swizzle atom
mma_f16(d_lo, dAlo + adv, dW + adv, (kb == 0 && kk == 0) ? 0u : 1u);
            mma_f16(d_lo, dA + adv, dWlo + adv, 1u);
            mma_f16(d_hi, dA + adv, dW + adv, (kb < NUM_HI && kk == 0) ? 0u : 1u);
          }
          umma_commit(&empty_bar[stage]);                          // frees the smem stage when the MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[0]);                                // accumulators complete
        acc_phase ^= 1;
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue =====================
    const int q = warp & 3;                                        // TMEM lane quarter this warp may access
    const int cb0 = ((warp - EPI_WARP0) >> 2) * CB_PER_WARP;
    uint32_t acc_phase = 0;
    const int nhi = p.k_blocks < NUM_HI ? p.k_blocks : NUM_HI;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int te = p.reverse ? p.num_tiles - 1 - tile : tile;
        const int tn = te % p.tiles_n, tm = te / p.tiles_n;
      const int g = tm / p.tiles_per_group, tb = tm - g * p.tiles_per_group;
      const int r_in_group = tb * BM + q * 32 + lane;
      const bool row_ok = r_in_group < p.rows_per_group;
      const long long m = (long long)g * p.rows_per_group + r_in_group;
      mbar_wait(&tfull_bar[0], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      epilogue_tile(p, tmem_base, q, cb0, tn, m, row_ok, nhi, stage, [&] {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[0]);
      });
      acc_phase ^= 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}


// --------------------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): a cluster of two CTAs computes a 256 x 128 output tile.  CTA r stages its own
// 128 rows of A (hi, lo: 2 x 16 KB) and HALF of the W tile (rows 64 r .. 64 r + 63 of the 128: 2 x 8 KB) per k-block;
// the leader CTA's MMA issuer runs tcgen05.mma.cta_group::2 with M = 256, N = 128: each SM's tensor core reads its
// own A tile and both halves of W, accumulators land in each CTA's own TMEM.  Shared-memory traffic per k-block and
// CTA drops from 160 KB (64 written by TMA + 96 read by the 12 MMAs) to 120 KB (48 + 72): the main loop of the
// single-CTA kernel is bound by exactly that traffic (128 B/clk per SM against 768 MMA cycles per k-block).
// Barriers: the leader's full barrier counts the bytes of both CTAs' TMA loads; MMA completion is committed by
// multicast to both CTAs' empty / accumulator-full barriers; both CTAs' epilogue warps arrive on the leader's
// accumulator-empty barrier.
constexpr int STAGES2 = 4;
constexpr int TILE_W2_BYTES = (BN / 2) * BK * 2;                 // 8 KB
constexpr int STAGE2_BYTES = 2 * TILE_BYTES + 2 * TILE_W2_BYTES;  // 48 KB
constexpr uint32_t kIdesc2 = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma2_load_3d(const CUtensorMap* map, uint32_t bar_cluster, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(const CUtensorMap* map, uint32_t bar_cluster, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mma2_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(kIdesc2), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar) {   // arrive on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_tc16x2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
                   const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapWlo,
                 const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(base + STAGES2 * STAGE2_BYTES);   // the leader's are the ones in use
  uint64_t* empty_bar = full_bar + STAGES2;
  uint64_t* tfull_bar = empty_bar + STAGES2;
  uint64_t* tempty_bar = tfull_bar + 1;                               // the leader's is the one in use
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* stage = (float*)(base + STAGES2 * STAGE2_BYTES + BAR_BYTES) + (warp >= EPI_WARP0 ? warp - EPI_WARP0 : 0) * STAGE_FLOATS_PER_WARP;
  const uint32_t rank = cluster_ctarank();
  const int pairs_per_group = (p.tiles_per_group + 1) / 2;
  const int num_pairs = p.tiles_n * pairs_per_group * p.n_groups;
  const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAlo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapWlo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES2; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar[0], 1); mbar_init(&tempty_bar[0], 2 * EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = cluster_id; t < num_pairs; t += n_clusters) {
        const int tn = t % p.tiles_n, tm = t / p.tiles_n;
        const int g = tm / pairs_per_group, pb = tm - g * pairs_per_group;
        const int tb = 2 * pb + (int)rank;                         // my 128-row tile inside the group
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = base + stage * STAGE2_BYTES;
          if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE2_BYTES);
          const uint32_t fb = mapa_rank(smem_u32(&full_bar[stage]), 0);
          tma2_load_3d(&mapA, fb, st, kb * BK, tb * BM, g);
          tma2_load_3d(&mapAlo, fb, st + TILE_BYTES, kb * BK, tb * BM, g);
          tma2_load_2d(&mapW, fb, st + 2 * TILE_BYTES, kb * BK, tn * BN + (int)rank * (BN / 2));
          tma2_load_2d(&mapWlo, fb, st + 2 * TILE_BYTES + TILE_W2_BYTES, kb * BK, tn * BN + (int)rank * (BN / 2));
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0;
      uint32_t acc_phase = 0;
      for (int t = cluster_id; t < num_pairs; t += n_clusters) {
        mbar_wait(&tempty_bar[0], acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d_lo = tmem_base + NUM_HI * BN;
        for (int kb = 0; kb < p.k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_u32(base + stage * STAGE2_BYTES);
          const uint64_t dA = make_desc(sa), dAlo = make_desc(sa + TILE_BYTES);
          const uint64_t dW = make_desc(sa + 2 * TILE_BYTES), dWlo = make_desc(sa + 2 * TILE_BYTES + TILE_W2_BYTES);
          const uint32_t d_hi = tmem_base + (uint32_t)(kb % NUM_HI) * BN;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t adv = (uint64_t)((kk * 32) >> 4);
            mma2_f16(d_lo, dAlo + adv, dW + adv, (kb == 0 && kk == 0) ? 0u : 1u);
            mma2_f16(d_lo, dA + adv, dWlo + adv, 1u);
            mma2_f16(d_hi, dA + adv, dW + adv, (kb < NUM_HI && kk == 0) ? 0u : 1u);
          }
          umma2_commit_mc(&empty_bar[stage]);                      // frees the stage in BOTH CTAs
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        umma2_commit_mc(&tfull_bar[0]);                            // accumulators complete, both CTAs
        acc_phase ^= 1;
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===================== epilogue (both CTAs, each its own 128 rows) =====================
    const int q = warp & 3;
    const int cb0 = ((warp - EPI_WARP0) >> 2) * CB_PER_WARP;
    uint32_t acc_phase = 0;
    const int nhi = p.k_blocks < NUM_HI ? p.k_blocks : NUM_HI;
    const uint32_t tempty0 = mapa_rank(smem_u32(&tempty_bar[0]), 0);
    for (int t = cluster_id; t < num_pairs; t += n_clusters) {
      const int tn = t % p.tiles_n, tm = t / p.tiles_n;
      const int g = tm / pairs_per_group, pb = tm - g * pairs_per_group;
      const int tb = 2 * pb + (int)rank;
      const int r_in_group = tb * BM + q * 32 + lane;
      const bool row_ok = r_in_group < p.rows_per_group;
      const long long m = (long long)g * p.rows_per_group + r_in_group;
      mbar_wait(&tfull_bar[0], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      epilogue_tile(p, tmem_base, q, cb0, tn, m, row_ok, nhi, stage, [&] {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty0);
      });
      acc_phase ^= 1;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();                                              // nobody leaves while the peer may still signal it
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// --------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)p;
  }
  return fn;
}

static bool encode(CUtensorMap* map, const __half* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box) {
  EncodeFn fn = get_encode();
  if (!fn) return false;
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, (void*)ptr, dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace tc16

static int g_tc16_pair = 0;       // 1: CTA-pair kernel (cta_group::2) where the launch has at least two 256-row tiles
void set_gemm_tc16_pair(int on) { g_tc16_pair = on ? 1 : 0; }

int launch_gemm_tc16(const Gemm16Args& g, cudaStream_t st) {
  using namespace tc16;
  if (!g.A_hi || !g.A_lo || !g.W_hi || !g.W_lo) return -100;
  if (g.Nout < 16 || g.K < BK || g.M < BM) return -100;
  if ((g.lda % 8) || (g.ldw % 8) || (g.a_group_stride % 8)) return -100;              // 16-byte row strides
  if (((uintptr_t)g.A_hi & 15) || ((uintptr_t)g.A_lo & 15) || ((uintptr_t)g.W_hi & 15) || ((uintptr_t)g.W_lo & 15)) return -100;
  if (g.C16_hi && ((g.ldc16 % 8) || ((uintptr_t)g.C16_hi & 15) || ((uintptr_t)g.C16_lo & 15) || !g.C16_lo)) return -100;
  Params p;
  const bool flat = g.a_rows_per_group >= g.M + g.row0;
  if (flat) {
    p.rows_per_group = g.M; p.n_groups = 1;
  } else {
    if (g.row0 % g.a_rows_per_group || g.M % g.a_rows_per_group) return -100;
    p.rows_per_group = g.a_rows_per_group; p.n_groups = g.M / g.a_rows_per_group;
  }
  p.tiles_per_group = (p.rows_per_group + BM - 1) / BM;
  p.M = g.M; p.Nout = g.Nout; p.K = g.K;
  p.tiles_n = (g.Nout + BN - 1) / BN;
  p.num_tiles = p.tiles_n * p.tiles_per_group * p.n_groups;
  p.k_blocks = (g.K + BK - 1) / BK;
  p.bias = g.bias; p.act = g.act;
  p.C16_hi = g.C16_hi; p.C16_lo = g.C16_lo; p.ldc16 = g.ldc16;
  p.C = g.C; p.ldc = g.ldc; p.n_store = g.n_store < g.Nout ? g.n_store : g.Nout;
  p.std32 = g.std32; p.mean32 = g.mean32; p.stat_rows_per_group = g.stat_rows_per_group; p.stat_ld = g.stat_ld;
  p.row0 = g.stat_row0; p.stat_mod = g.stat_mod;
  p.overflow = g.overflow_flag;
  p.reverse = g.reverse;

  const long long a_off = flat ? (long long)g.row0 * g.lda
                               : (long long)(g.row0 / g.a_rows_per_group) * g.a_group_stride;
  const long long gstride = flat ? (long long)g.M * g.lda : g.a_group_stride;
  CUtensorMap mA, mAlo, mW, mWlo;
  {
    cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)p.rows_per_group, (cuuint64_t)p.n_groups};
    cuuint64_t strides[2] = {(cuuint64_t)g.lda * 2, (cuuint64_t)gstride * 2};
    cuuint32_t box[3] = {BK, BM, 1};
    if (!encode(&mA, g.A_hi + a_off, 3, dims, strides, box)) return -100;
    if (!encode(&mAlo, g.A_lo + a_off, 3, dims, strides, box)) return -100;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)g.K, (cuuint64_t)g.Nout};
    cuuint64_t strides[1] = {(cuuint64_t)g.ldw * 2};
    const int pairs = p.tiles_n * ((p.tiles_per_group + 1) / 2) * p.n_groups;
    const bool use_pair = g_tc16_pair && pairs >= 2;
    cuuint32_t box[2] = {BK, (cuuint32_t)(use_pair ? BN / 2 : BN)};          // CTA pair: each CTA stages half of the W tile
    if (!encode(&mW, g.W_hi, 2, dims, strides, box)) return -100;
    if (!encode(&mWlo, g.W_lo, 2, dims, strides, box)) return -100;
    if (use_pair) {
      const size_t smem2 = (size_t)STAGES2 * STAGE2_BYTES + 1024 + BAR_BYTES + EPI_STAGE_BYTES;
      static PerDeviceInt sm_table2;
      const int sm2 = sm_table2.get([&] {
        int dev = 0, n = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(gemm_tc16x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        return n;
      });
      int clusters = pairs < sm2 / 2 ? pairs : sm2 / 2;
      gemm_tc16x2_kernel<<<2 * clusters, NUM_THREADS, smem2, st>>>(mA, mAlo, mW, mWlo, p);
      return (int)cudaGetLastError();
    }
  }
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024 + BAR_BYTES + EPI_STAGE_BYTES;
  static PerDeviceInt sm_table;           // SM count; the shared-memory attribute is set on the same first use
  const int sm_count = sm_table.get([&] {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(gemm_tc16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return n;
  });
  int grid = p.num_tiles < sm_count ? p.num_tiles : sm_count;
  gemm_tc16_kernel<<<grid, NUM_THREADS, smem, st>>>(mA, mAlo, mW, mWlo, p);
  return (int)cudaGetLastError();
}

// fp32 matrix [rows, cols] (row stride ld_in) -> fp16 pair (hi, lo * 2^11), row stride ld_out >= cols, padding zeroed
__global__ void split16_kernel(const float* __restrict__ x, long long rows, int cols, int ld_in, __half* __restrict__ hi,
                               __half* __restrict__ lo, int ld_out, int* __restrict__ overflow) {
  const long long total = rows * ld_out;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / ld_out;
    const int c = (int)(i - r * ld_out);
    float v = 0.0f;
    if (c < cols) v = x[r * ld_in + c];
    const __half h = __float2half_rn(v);
    hi[i] = h;
    lo[i] = __float2half_rn((v - __half2float(h)) * 2048.0f);
    if (overflow && !(fabsf(v) <= 65504.0f)) *overflow = 1;
  }
}
int launch_split16(const float* x, long long rows, int cols, int ld_in, __half* hi, __half* lo, int ld_out, int* overflow,
                   cudaStream_t st) {
  const long long n = rows * ld_out;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  split16_kernel<<<(int)blocks, 256, 0, st>>>(x, rows, cols, ld_in, hi, lo, ld_out, overflow);
  return (int)cudaGetLastError();
}

}  // namespace kmpc
