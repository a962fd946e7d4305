// tcgen05 / TMA GEMM with fp32 accuracy (3xTF32 operand split) for the bulk layers of the forecast path:
//     C[M, Nout] = epilogue( A[M,K] . W[Nout,K]^T ),   A, W fp32 K-major (nn.Linear layout, model.py:96-117)
//
// Every fp32 operand x is used as x = x_hi + x_lo with x_hi = the top 19 bits of x (what kind::tf32 reads from an
// fp32 container: the low 13 mantissa bits are ignored) and x_lo = x - x_hi (exact in fp32), kept in a twin
// buffer.  D = A_hi.W_hi + A_lo.W_hi + A_hi.W_lo accumulated in fp32 in TMEM; the dropped A_lo.W_lo term and the
// truncation of x_lo are O(2^-20) relative.  The epilogue writes C and its own C_lo twin for the next layer.
//
// TMEM accumulation truncates the running sum at every tcgen05.mma (measured: the error of a single accumulator
// grows linearly with the number of accumulate steps, 1.2e-5 at K=1024 with 384 steps).  The tile therefore uses
// four accumulators: the hi.hi products of k-block kb go to accumulator kb % 3, every lo product goes to a fourth
// one (its magnitude is 2^-10 of the result, so its truncation is invisible); the epilogue adds the four in fp32
// with round-to-nearest.  512 TMEM columns = one tile in flight.
//
// Structure (one CTA per SM, persistent over output tiles, 128 x 128 x 32 tiles, 3-stage TMA ring):
//   warp 0      TMA producer: 4 boxes per stage (A, A_lo, W, W_lo), SWIZZLE_128B, mbarrier complete_tx
//   warp 1      MMA issuer: one elected lane, 12 tcgen05.mma.kind::tf32 per stage, tcgen05.commit to free the stage
//   warp 2      TMEM allocator (4 accumulators x 128 columns)
//   warps 4..7  epilogue: tcgen05.ld 32x32b.x32 -> bias / addend / activation / shrink -> global stores
// The A operand is addressed through a 3-D tensor map (k, row-in-group, group) so that the delay-embedded window
// view of the standardised series (rows overlap, never materialised) and plain activation matrices use the same
// kernel.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "gemm.cuh"

namespace kmpc {

namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;          // BK fp32 = 128 bytes = one swizzle atom row
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;             // 16 KB (BM == BN)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;         // A, A_lo, W, W_lo
constexpr int NUM_THREADS = 256;
constexpr int EPI_WARP0 = 4;
constexpr int TMEM_COLS = 512;                      // 3 hi accumulators + 1 lo accumulator of 128 fp32 columns
constexpr int NUM_HI = 3;
constexpr uint32_t SPIN_LIMIT = 1u << 28;           // bounded waits: trap instead of hanging the GPU

struct Params {
  int M, Nout, K;
  int rows_per_group, tiles_per_group, n_groups;     // M = n_groups * rows_per_group
  int tiles_n, num_tiles, k_blocks;
  const float* bias;
  const float* addend; int ld_add;
  int act; float shrink_thr;
  const float* std32; const float* mean32; int stat_rows_per_group; int stat_ld; int row0; int stat_mod;   // de-standardise epilogue
  float* C; float* C_lo; long long ldc; int n_store;
  const int* gate;
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
    if (++spins > SPIN_LIMIT) { printf("kmpc gemm_tc: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
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
// instruction descriptor (UMMA::InstrDescriptor): c=F32, a=b=TF32, K-major both, N=128, M=128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
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

// Interior block (16 full columns, 16-byte aligned rows, no de-standardisation): bias / addend as vector loads,
// activation resolved at compile time, C and its residual twin as float4 stores.  The per-element general path
// below (branches per element) had made the epilogue the bottleneck of the whole GEMM (32 us of a 58 us tile).
template <int ACT>
__device__ __forceinline__ void store_block_fast(const Params& p, const float (&acc)[16], long long m, int n0) {
  float x[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = acc[j];
  if (p.bias) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) { const float4 b = b4[j4]; x[j4 * 4] += b.x; x[j4 * 4 + 1] += b.y; x[j4 * 4 + 2] += b.z; x[j4 * 4 + 3] += b.w; }
  }
  if (p.addend) {
    const float4* a4 = reinterpret_cast<const float4*>(p.addend + m * p.ld_add + n0);
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) { const float4 b = a4[j4]; x[j4 * 4] += b.x; x[j4 * 4 + 1] += b.y; x[j4 * 4 + 2] += b.z; x[j4 * 4 + 3] += b.w; }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) x[j] = epilogue_apply(x[j], ACT, p.shrink_thr);
  float4* crow = reinterpret_cast<float4*>(p.C + m * p.ldc + n0);
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) crow[j4] = make_float4(x[j4 * 4], x[j4 * 4 + 1], x[j4 * 4 + 2], x[j4 * 4 + 3]);
  if (p.C_lo) {
    float4* lrow = reinterpret_cast<float4*>(p.C_lo + m * p.ldc + n0);
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      float lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float t = x[j4 * 4 + j]; lo[j] = tf32_residual(t); }
      lrow[j4] = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapWlo, Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: stages first (1024-aligned), then barriers
  uint8_t* base = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = (uint64_t*)(base + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;     // [1] accumulators ready
  uint64_t* tempty_bar = tfull_bar + 1;         // [1] accumulators drained
  uint32_t* tmem_slot = (uint32_t*)(tempty_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (p.gate && *p.gate == 0) return;           // gated launch (see GemmArgs::gate): uniform over the grid

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapAlo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapWlo) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar[0], 1); mbar_init(&tempty_bar[0], 4);
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
        const int tn = tile % p.tiles_n, tm = tile / p.tiles_n;
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
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t adv = (uint64_t)((kk * 32) >> 4);       // 8 fp32 = 32 bytes inside the swizzle atom
            mma_tf32(d_lo, dAlo + adv, dW + adv, (kb == 0 && kk == 0) ? 0u : 1u);
            mma_tf32(d_lo, dA + adv, dWlo + adv, 1u);
            mma_tf32(d_hi, dA + adv, dW + adv, (kb < NUM_HI && kk == 0) ? 0u : 1u);
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
    uint32_t acc_phase = 0;
    const int nhi = p.k_blocks < NUM_HI ? p.k_blocks : NUM_HI;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int tn = tile % p.tiles_n, tm = tile / p.tiles_n;
      const int g = tm / p.tiles_per_group, tb = tm - g * p.tiles_per_group;
      const int r_in_group = tb * BM + q * 32 + lane;
      const bool row_ok = r_in_group < p.rows_per_group;
      const long long m = (long long)g * p.rows_per_group + r_in_group;
      mbar_wait(&tfull_bar[0], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      // vector path: rows of C / C_lo / addend 16-byte aligned, no de-standardisation
      const bool fast_ok = !p.std32 && (p.ldc % 4 == 0) && ((((uintptr_t)p.C) & 15) == 0) &&
                           (!p.C_lo || (((uintptr_t)p.C_lo) & 15) == 0) &&
                           (!p.addend || ((p.ld_add % 4 == 0) && (((uintptr_t)p.addend) & 15) == 0)) &&
                           (!p.bias || (((uintptr_t)p.bias) & 15) == 0);
#pragma unroll 1
      for (int cb = 0; cb < BN / 16; ++cb) {
        uint32_t v[16], u0[16], u1[16], u2[16];
        float acc[16];
        tmem_ld16_nowait(lane_addr + (uint32_t)(NUM_HI * BN + cb * 16), v);    // lo products
        tmem_ld16_nowait(lane_addr + (uint32_t)(0 * BN + cb * 16), u0);
        if (nhi > 1) tmem_ld16_nowait(lane_addr + (uint32_t)(1 * BN + cb * 16), u1);
        if (nhi > 2) tmem_ld16_nowait(lane_addr + (uint32_t)(2 * BN + cb * 16), u2);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float a = __fadd_rn(__uint_as_float(v[j]), __uint_as_float(u0[j]));
          if (nhi > 1) a = __fadd_rn(a, __uint_as_float(u1[j]));
          if (nhi > 2) a = __fadd_rn(a, __uint_as_float(u2[j]));
          acc[j] = a;
        }
        const int n0 = tn * BN + cb * 16;
        if (row_ok && fast_ok && n0 + 16 <= p.n_store) {
          switch (p.act) {
            case EPI_RELU: store_block_fast<EPI_RELU>(p, acc, m, n0); break;
            case EPI_TANH: store_block_fast<EPI_TANH>(p, acc, m, n0); break;
            case EPI_GELU: store_block_fast<EPI_GELU>(p, acc, m, n0); break;
            case EPI_SHRINK: store_block_fast<EPI_SHRINK>(p, acc, m, n0); break;
            default: store_block_fast<EPI_NONE>(p, acc, m, n0); break;
          }
        } else if (row_ok && n0 < p.n_store) {
          float* crow = p.C + m * p.ldc + n0;
          float* lrow = p.C_lo ? p.C_lo + m * p.ldc + n0 : nullptr;
          const float* arow = p.addend ? p.addend + m * p.ld_add + n0 : nullptr;
          const long long sg = (p.std32 && p.stat_rows_per_group > 0) ? (m + p.row0) / p.stat_rows_per_group : 0;
          const bool full = (n0 + 16 <= p.n_store) && ((((uintptr_t)crow) & 15) == 0);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            float x[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int n = n0 + j4 * 4 + j;
              float t = acc[j4 * 4 + j];
              if (n < p.n_store) {
                if (p.bias) t += p.bias[n];
                if (arow) t += arow[j4 * 4 + j];
              }
              t = epilogue_apply(t, p.act, p.shrink_thr);
              if (p.std32 && n < p.n_store) {
                const int sn = p.stat_mod ? n % p.stat_mod : n;
                t = __fadd_rn(__fmul_rn(t, p.std32[sg * p.stat_ld + sn]), p.mean32[sg * p.stat_ld + sn]);
              }
              x[j] = t;
              lo[j] = tf32_residual(t);
            }
            if (full) {
              *reinterpret_cast<float4*>(crow + j4 * 4) = make_float4(x[0], x[1], x[2], x[3]);
              if (lrow) *reinterpret_cast<float4*>(lrow + j4 * 4) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (n0 + j4 * 4 + j < p.n_store) { crow[j4 * 4 + j] = x[j]; if (lrow) lrow[j4 * 4 + j] = lo[j]; }
            }
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[0]);
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

static bool encode(CUtensorMap* map, const float* ptr, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box) {
  EncodeFn fn = get_encode();
  if (!fn) return false;
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, (void*)ptr, dims, strides_bytes, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace tc

static int g_tc_mode = 1;      // 1 = use the tcgen05 path when eligible, 0 = force the SIMT kernel (diagnostics)
void set_gemm_tc_mode(int on) { g_tc_mode = on; }

int launch_gemm_tc(const GemmArgs& g, cudaStream_t st) {
  using namespace tc;
  if (!g_tc_mode) return -100;
  if (!g.A_lo || !g.W_lo) return -100;
  if (g.Nout < 16 || g.K < BK || g.M < BM) return -100;
  if ((g.lda % 4) || (g.ldw % 4) || (g.a_group_stride % 4)) return -100;
  if (((uintptr_t)g.A & 15) || ((uintptr_t)g.A_lo & 15) || ((uintptr_t)g.W & 15) || ((uintptr_t)g.W_lo & 15)) return -100;
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
  p.bias = g.bias; p.addend = g.addend; p.ld_add = g.ld_add; p.act = g.act; p.shrink_thr = g.shrink_thr;
  p.C = g.C; p.C_lo = g.C_lo; p.ldc = g.ldc; p.n_store = g.n_store < g.Nout ? g.n_store : g.Nout;
  p.std32 = g.std32; p.mean32 = g.mean32; p.stat_rows_per_group = g.stat_rows_per_group; p.stat_ld = g.stat_ld; p.row0 = g.stat_row0; p.stat_mod = g.stat_mod;
  p.gate = g.gate;

  const long long a_off = flat ? (long long)g.row0 * g.lda
                               : (long long)(g.row0 / g.a_rows_per_group) * g.a_group_stride;
  const long long gstride = flat ? (long long)g.M * g.lda : g.a_group_stride;
  CUtensorMap mA, mAlo, mW, mWlo;
  {
    cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)p.rows_per_group, (cuuint64_t)p.n_groups};
    cuuint64_t strides[2] = {(cuuint64_t)g.lda * 4, (cuuint64_t)gstride * 4};
    cuuint32_t box[3] = {BK, BM, 1};
    if (!encode(&mA, g.A + a_off, 3, dims, strides, box)) return -100;
    if (!encode(&mAlo, g.A_lo + a_off, 3, dims, strides, box)) return -100;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)g.K, (cuuint64_t)g.Nout};
    cuuint64_t strides[1] = {(cuuint64_t)g.ldw * 4};
    cuuint32_t box[2] = {BK, BN};
    if (!encode(&mW, g.W, 2, dims, strides, box)) return -100;
    if (!encode(&mWlo, g.W_lo, 2, dims, strides, box)) return -100;
  }
  const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024 + 256;
  static PerDeviceInt sm_table;           // SM count; the shared-memory attribute is set on the same first use
  const int sm_count = sm_table.get([&] {
    int dev = 0, n = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    return n;
  });
  int grid = p.num_tiles < sm_count ? p.num_tiles : sm_count;
  gemm_tc_kernel<<<grid, NUM_THREADS, smem, st>>>(mA, mAlo, mW, mWlo, p);
  return (int)cudaGetLastError();
}

// x_lo = x - top19bits(x) for a whole buffer (weights at load time, the standardised series per forecast call)
__global__ void split_lo_kernel(const float* __restrict__ x, float* __restrict__ lo, long long n, const int* __restrict__ gate) {
  if (gate && *gate == 0) return;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float v = x[i];
    lo[i] = tf32_residual(v);
  }
}
int launch_split_lo(const float* x, float* lo, long long n, cudaStream_t st, const int* gate) {
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  split_lo_kernel<<<(int)blocks, 256, 0, st>>>(x, lo, n, gate);
  return (int)cudaGetLastError();
}

}  // namespace kmpc
