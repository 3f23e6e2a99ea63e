// Round-2 probe (NOT on the product path): one 128 x 128 x 64 tcgen05 kind::f16 MMA tile whose B operand is MN-major
// (N contiguous, as a row-major [K][N] fp16 matrix is) with the shared-memory descriptor fields supplied at run time, so
// that one GPU call can sweep candidate (LBO, SBO, layout, K-step) encodings and find the one the hardware takes
// (tools/probe_mn16.py).  If 16-bit MN-major operands work, the node-transposed copies (X16T, dV16T, Q16T) and their
// epilogue stores can be dropped: the fused kernels would read the row-major copies directly.
//   A: fp16 [128][64] K-major (box [64][128], SWIZZLE_128B)       B: fp16 [64 (k)][128 (n)] (two boxes [64 n][64 k], SWIZZLE_128B)
#pragma once

#include "agcn_fused_h.cuh"

namespace mcrn {
namespace probe {

using namespace tc;

struct ProbeParams {
  uint32_t lbo, sbo, layout, kstep;   // descriptor fields of B; kstep = byte advance of the start address per K = 16 step
  uint32_t b_major;                   // instruction-descriptor b_major bit (1 = MN-major)
  float* C;                           // [128][128]
};

__global__ void __launch_bounds__(128, 1)
probe_mn16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar, done_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_addr = smem_base, b_addr = smem_base + 16384;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&full_bar), 1);
    mbar_init(smem_u32(&done_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(smem_u32(&full_bar), 16384 + 16384);
    tma_load_4d(a_addr, &tmA, smem_u32(&full_bar), 0, 0, 0, 0);              // A[0..128][0..64]
    tma_load_4d(b_addr, &tmB, smem_u32(&full_bar), 0, 0, 0, 0);              // B[0..64][n 0..64]
    tma_load_4d(b_addr + 8192, &tmB, smem_u32(&full_bar), 64, 0, 0, 0);      // B[0..64][n 64..128]
    fused::mbar_wait_b(smem_u32(&full_bar), 0);
    tcgen05_fence_after();
    const uint32_t idesc = (1u << 4) | ((p.b_major & 1u) << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int kk = 0; kk < 4; ++kk) {
      const uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024, 2);
      const uint64_t bd = make_smem_desc(b_addr + kk * p.kstep, p.lbo, p.sbo, p.layout);
      fusedh::tcgen05_mma_f16(tmem_base, ad, bd, idesc, kk > 0 ? 1u : 0u);
    }
    tcgen05_commit(smem_u32(&done_bar));
  }
  fused::mbar_wait_b(smem_u32(&done_bar), 0);
  tcgen05_fence_after();
  float v[32];
  for (int c = 0; c < 4; ++c) {
    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 32), v);
    for (int j = 0; j < 32; ++j) p.C[(warp * 32 + lane) * 128 + c * 32 + j] = v[j];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128) : "memory");
}

// A: device fp16 [128][64]; B: device fp16 [64][128]; C: device fp32 [128][128]
static inline int run_probe_mn16(const __half* A, const __half* B, float* C, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t kstep,
                                 uint32_t b_major, cudaStream_t st) {
  CUtensorMap tA, tB;
  {
    uint64_t dims[4] = {64, 128, 1, 1};
    uint64_t str[3] = {64 * 2, 128 * 64 * 2, 128 * 64 * 2};
    uint32_t box[4] = {64, 128, 1, 1};
    MCRN_TRY(fusedh::encode_tensor_map_h(&tA, A, dims, str, box));
  }
  {
    uint64_t dims[4] = {128, 64, 1, 1};
    uint64_t str[3] = {128 * 2, 64 * 128 * 2, 64 * 128 * 2};
    uint32_t box[4] = {64, 64, 1, 1};
    MCRN_TRY(fusedh::encode_tensor_map_h(&tB, B, dims, str, box));
  }
  ProbeParams p{lbo, sbo, layout, kstep, b_major, C};
  static bool attr_set = false;
  if (!attr_set) {
    MCRN_CUDA_OK(cudaFuncSetAttribute(probe_mn16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024));
    attr_set = true;
  }
  MCRN_LAUNCH(probe_mn16_kernel, 1, 128, 34 * 1024, st, tA, tB, p);
  return MCRN_OK;
}

}  // namespace probe
}  // namespace mcrn
