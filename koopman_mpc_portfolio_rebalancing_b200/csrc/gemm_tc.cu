// tcgen05 / TMA 3xTF32 GEMM (placeholder: reports "not eligible" until the kernel lands)
#include "gemm.cuh"
namespace kmpc { int launch_gemm_tc(const GemmArgs&, cudaStream_t) { return -100; } }
