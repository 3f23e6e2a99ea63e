// tcgen05 TF32 GEMM instantiations (placeholder translation unit until the kernel lands).
#include "common.cuh"
