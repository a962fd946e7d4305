// Microbenchmark: throughput and dependent-issue latency of mma.sync.m8n8k4.f64 (DMMA) vs DFMA on one SM and on the
// whole chip.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_bench dmma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int CHAINS>
__global__ void k_dmma(double* out, int iters, long long* cycles) {
  double c0[CHAINS], c1[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) dmma(c0[i], c1[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int CHAINS>
__global__ void k_dfma(double* out, int iters, long long* cycles) {
  double c[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) c[i] = threadIdx.x * 1e-3 + i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i] = fma(c[i], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

__global__ void k_shfl(double* out, int iters, long long* cycles) {
  double v = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) v = fma(__shfl_sync(0xffffffffu, v, (it + 1) & 31), 1.0000001, v);
  long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}
__global__ void k_rcp(double* out, int iters, long long* cycles) {
  double x = 1.5 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    x = fma(y, fma(e, e, e), y) + 1.25;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}
__global__ void k_lds(double* out, int iters, long long* cycles) {
  __shared__ double s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 37) % 1024;
  __syncthreads();
  int idx = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) idx = (int)s[idx];
  long long t1 = clock64();
  out[threadIdx.x] = idx;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

template <typename F>
void run(const char* name, F launch, int iters, double ops_per_thread_iter, int blocks, int threads) {
  double* out; long long* cyc; cudaMalloc(&out, sizeof(double) * blocks * threads); cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(out, iters, cyc); cudaDeviceSynchronize();
  cudaEventRecord(e0); launch(out, iters, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s blocks %4d thr %4d: %8.1f cyc/iter (block 0), %.3f ms, %.2f Tops/s  err=%s\n", name, blocks, threads,
         (double)c / iters, ms, ops_per_thread_iter * iters * blocks * threads / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  const int it = 20000;
  // latency: 1 warp, 1 dependent chain
  run("DMMA 1 chain, 1 warp (latency)", [](double* o, int i, long long* c) { k_dmma<1><<<1, 32>>>(o, i, c); }, it, 2 * 256.0 / 32, 1, 32);
  run("DMMA 6 chains, 1 warp", [](double* o, int i, long long* c) { k_dmma<6><<<1, 32>>>(o, i, c); }, it, 6 * 2 * 256.0 / 32, 1, 32);
  run("DMMA 6 chains, 4 warps (1 per SMSP)", [](double* o, int i, long long* c) { k_dmma<6><<<1, 128>>>(o, i, c); }, it, 6 * 2 * 256.0 / 32, 1, 128);
  run("DMMA 6 chains, 8 warps", [](double* o, int i, long long* c) { k_dmma<6><<<1, 256>>>(o, i, c); }, it, 6 * 2 * 256.0 / 32, 1, 256);
  run("DMMA 6 chains, 8 warps x 148 blocks", [](double* o, int i, long long* c) { k_dmma<6><<<148, 256>>>(o, i, c); }, it, 6 * 2 * 256.0 / 32, 148, 256);
  run("DFMA 1 chain, 1 warp (latency)", [](double* o, int i, long long* c) { k_dfma<1><<<1, 32>>>(o, i, c); }, it, 2, 1, 32);
  run("DFMA 8 chains, 1 warp", [](double* o, int i, long long* c) { k_dfma<8><<<1, 32>>>(o, i, c); }, it, 16, 1, 32);
  run("DFMA 8 chains, 4 warps", [](double* o, int i, long long* c) { k_dfma<8><<<1, 128>>>(o, i, c); }, it, 16, 1, 128);
  run("DFMA 8 chains, 8 warps", [](double* o, int i, long long* c) { k_dfma<8><<<1, 256>>>(o, i, c); }, it, 16, 1, 256);
  run("DFMA 8 chains, 8 warps x 148 blocks", [](double* o, int i, long long* c) { k_dfma<8><<<148, 256>>>(o, i, c); }, it, 16, 148, 256);
  run("SHFL.f64 + DFMA dependent chain", [](double* o, int i, long long* c) { k_shfl<<<1, 32>>>(o, i, c); }, it, 1, 1, 32);
  run("rcp_fast (MUFU.RCP64H + 3 DFMA) + DADD chain", [](double* o, int i, long long* c) { k_rcp<<<1, 32>>>(o, i, c); }, it, 1, 1, 32);
  run("LDS.64 dependent chain (+I2F/F2I)", [](double* o, int i, long long* c) { k_lds<<<1, 32>>>(o, i, c); }, it, 1, 1, 32);
  return 0;
}
