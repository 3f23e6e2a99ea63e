// GEMM engine dispatch: tcgen05 TF32 (gemm_tc.cuh) when the problem meets the TMA/UMMA
// constraints, SIMT fp32 (gemm_simt.cuh) otherwise or when forced with mcrn_set_engine(1).
#pragma once

#include "gemm_simt.cuh"
#include "gemm_tc.cuh"

namespace mcrn {

extern int g_engine;   // 0 = default, 1 = force SIMT, 2 = force tcgen05 (error if not eligible)

template <class Epi>
int gemm(const GemmDesc& q, const Epi& e, cudaStream_t st) {
  if (g_engine != 1 && tc::eligible(q)) return q.b_sub == 2 ? tc::gemm_tc_hilo(q, e, st) : tc::gemm_tc(q, e, st);
  if (g_engine == 2) {
    set_error("engine 2 (tcgen05) forced but the problem is not eligible (alignment / strides)");
    return MCRN_ERR_BAD_DIMS;
  }
  return gemm_simt(q, e, st);
}

}  // namespace mcrn
