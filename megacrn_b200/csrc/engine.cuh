// GEMM engine dispatch: tcgen05 TF32 (gemm_tc.cuh) when the problem meets the TMA/UMMA
// constraints, SIMT fp32 (gemm_simt.cuh) otherwise or when forced with mcrn_set_engine(1).
#pragma once

#include "gemm_simt.cuh"

namespace mcrn {

extern int g_engine;   // 0 = default, 1 = force SIMT, 2 = force tcgen05 (error if not eligible)

template <class Epi>
int gemm(const GemmDesc& q, const Epi& e, cudaStream_t st) {
  return gemm_simt(q, e, st);
}

}  // namespace mcrn
