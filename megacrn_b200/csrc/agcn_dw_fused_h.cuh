// Weight gradients of ONE AGCN over all time steps and batch elements, fp16 operands (tcgen05 kind::f16, fp32 accumulate):
//     dW[blk][c][h*HS + o] = sum_{t,b,n} X_t[n,b,c] * Y_t[n,b,o]
// with Y = dV (blk = 0, the folded identity block) or Y = Q_{k,h} = S_k^T dV (blk = 1 + k; agcn_bwd_fused_h.cuh).  Both
// operands are the ROW-MAJOR fp16 copies the fused forward / backward write anyway, read as MN-major tcgen05 operands
// (the contraction index -- the node -- is the strided one; descriptor: LBO 8192 / SBO 1024 / SWIZZLE_128B, 2048 B per
// K = 16 step, a_major = b_major = 1; tools/probe_mn16.py), so no node-transposed copy exists anywhere:
//     X16  [T][R][HS]               A: 2 boxes [64 c][1][64 nodes] (the second is out of bounds = 0 when HS = 64)
//     dV16 [T][R][O]                B for blk 0: HS/64 boxes at column half * HS
//     Q16  [T * nks][R][HS]         B for blk >= 1: block ks = k * nhalf + h
//     IB16 [T][R][64]               A of the input-block tiles (blk = KS + 1: rows = input-channel / bias columns of the compact
//                                   input block, 16 used), with Y = dV: dW_NB = sum IB^T dV  (replaces a TF32 GEMM over fp32 dV)
// A plain split-K GEMM: CTA = (output tile (blk, h), group g) loops over its (t, b) units, ceil(N / 64) ring items each,
// accumulator [128 x HS] in TMEM, one red.global.add.v4.f32 pass at the end (times 1 / loss scale: dV16T and Q16T carry it).
// Warp roles: warps 0, 6, 7 = TMA producers (ring items round-robin), warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#pragma once

#include "agcn_bwd_fused_h.cuh"

namespace mcrn {
namespace fusedwh {

using namespace tc;
using fused::mbar_wait_b;
using fusedh::BKH;
using fusedh::make_idesc_f16;
using fusedh::tcgen05_mma_f16;

constexpr int NPROD_W = 3;                      // TMA producer warps (one thread issues ~1 bulk-tensor copy per 190 cycles; an item is 3-4 of them)
constexpr int WTHREADS = 192 + 32 * (NPROD_W - 1);

struct WParams {
  int N, B, T, KS, nhalf, O;
  float* dw;             // [KS+2][HS][O] accumulators
  const float* gs;       // device {scale, 1 / scale}
};

template <int HS>
struct CfgW {
  static constexpr uint32_t A_SLOT = BM * 128, B_SLOT = (uint32_t)HS * 128, STAGE = A_SLOT + B_SLOT;
  static constexpr int NST = 6;
  static constexpr size_t SMEM = (size_t)NST * STAGE + 1024;
  static constexpr uint32_t TMEM_COLS = HS >= 128 ? 128 : 64;
};

template <int HS>
__global__ void __launch_bounds__(WTHREADS, 1)
agcn_dw_h_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmIB, WParams p) {
  using C = CfgW<HS>;
  constexpr int NST = C::NST;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[NST];
  __shared__ __align__(8) uint64_t empty_bar[NST];
  __shared__ __align__(8) uint64_t acc_full_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, grp = blockIdx.y, G = gridDim.y;
  const int blk = tile / p.nhalf, half = tile - blk * p.nhalf;      // blk 0 = identity block, 1 + k = support k, KS + 1 = input block
  const bool ib_tile = blk == p.KS + 1;
  const int U = p.T * p.B;
  const int nu = (U - grp + G - 1) / G;
  const int kb = (p.N + BKH - 1) / BKH;
  const int nks = p.KS * p.nhalf;
  if (nu <= 0) return;
  const int nit = nu * kb;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmIB) : "memory");
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&acc_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(C::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_slot;

  const int pw = warp == 0 ? 0 : (warp >= 6 ? warp - 5 : -1);      // producer index or -1
  if (pw >= 0) {
    if (lane == 0) {                                     // ===== TMA producers: item `it` belongs to producer it % NPROD_W =====
      for (int it = pw; it < nit; it += NPROD_W) {
        const int s = it % NST;
        if (it >= NST) mbar_wait_b(smem_u32(&empty_bar[s]), (((uint32_t)(it / NST)) & 1u) ^ 1u);
        const uint32_t fb = smem_u32(&full_bar[s]);
        const uint32_t a_dst = smem_base + (uint32_t)s * C::STAGE, b_dst = a_dst + C::A_SLOT;
        const int i = it / kb, j = it - i * kb;
        const int u = grp + i * G, t = u / p.B, b = u - t * p.B;
        mbar_expect_tx(fb, C::A_SLOT + C::B_SLOT);
        const CUtensorMap* ta = ib_tile ? &tmIB : &tmX;
        tma_load_4d(a_dst, ta, fb, 0, b, j * BKH, t);                                     // X_t / IB_t rows (64 j.., b), channels 0..63
        tma_load_4d(a_dst + 8192, ta, fb, 64, b, j * BKH, t);                             // channels 64..127 (out of bounds = zero fill when the operand is 64 wide)
#pragma unroll
        for (int q = 0; q < HS / 64; ++q) {
          if (blk == 0 || ib_tile) tma_load_4d(b_dst + q * 8192, &tmV, fb, half * HS + q * 64, b, j * BKH, t);      // dV_t rows, columns half*HS + 64 q..
          else tma_load_4d(b_dst + q * 8192, &tmQ, fb, q * 64, b, j * BKH, t * nks + (blk - 1) * p.nhalf + half);   // Q_t[ks] rows
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                     // ===== MMA issuer =====
      constexpr uint32_t idesc = make_idesc_f16<HS>() | (1u << 15) | (1u << 16);     // A and B MN-major
      for (int it = 0; it < nit; ++it) {
        const int s = it % NST;
        mbar_wait_b(smem_u32(&full_bar[s]), ((uint32_t)(it / NST)) & 1u);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_base + (uint32_t)s * C::STAGE, b_addr = a_addr + C::A_SLOT;
#pragma unroll
        for (int kk = 0; kk < BKH / 16; ++kk) {
          const uint64_t ad = make_smem_desc(a_addr + kk * 2048, 8192, 1024, 2);
          const uint64_t bd = make_smem_desc(b_addr + kk * 2048, 8192, 1024, 2);
          tcgen05_mma_f16(tmem_base, ad, bd, idesc, (it > 0 || kk > 0) ? 1u : 0u);
        }
        tcgen05_commit(smem_u32(&empty_bar[s]));
      }
      tcgen05_commit(smem_u32(&acc_full_bar));
    }
  } else if (warp < 6) {                                 // ===== epilogue warps =====
    const int quarter = warp & 3;
    mbar_wait_b(smem_u32(&acc_full_bar), 0);
    tcgen05_fence_after();
    const int c0 = quarter * 32;                         // accumulator rows = weight rows c
    const int rows = ib_tile ? fusedh::IBF : HS;         // input-block tile: only the first 16 rows exist
    if (c0 < rows) {
      const float inv_gs = __ldg(p.gs + 1);
      float* scr = reinterpret_cast<float*>(smem_al) + (warp - 2) * (32 * 36);     // the ring is idle now
      const int cq = (lane & 7) * 4, r0 = lane >> 3;
      float* dst = p.dw + (int64_t)blk * HS * p.O + half * HS;
#pragma unroll 1
      for (int ch = 0; ch < HS / 32; ++ch) {
        float v[32];
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ch * 32), v);
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          *reinterpret_cast<float4*>(&scr[lane * 36 + e]) = make_float4(v[e] * inv_gs, v[e + 1] * inv_gs, v[e + 2] * inv_gs, v[e + 3] * inv_gs);
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int rr = r0 + 4 * e, c = c0 + rr;
          if (c < rows) {
            const float4 t4 = *reinterpret_cast<const float4*>(&scr[rr * 36 + cq]);
            atomicAdd(reinterpret_cast<float4*>(dst + (int64_t)c * p.O + ch * 32 + cq), t4);
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

// x16_all [T][R][HS], v16_all [T][R][O], q16_all [T * KS * nhalf][R][HS]: row-major, rows (node, b)  (steps [0, T) of the launch)
template <int HS>
int launch_agcn_dw_h(int N, int B, int T, int KS, int nhalf, const __half* x16_all, const __half* v16_all, const __half* q16_all,
                     const __half* ib16_all /* [T][R][64] compact input blocks, or null */, const float* gs, float* dw, cudaStream_t st) {
  using C = CfgW<HS>;
  const int O = nhalf * HS, nks = KS * nhalf;
  const uint64_t R = (uint64_t)N * B;
  CUtensorMap tX, tV, tQ, tIB;
  uint32_t box[4] = {64, 1, BKH, 1};                     // 64 channels x 64 node rows of one batch element
  {
    uint64_t dims[4] = {(uint64_t)fusedh::IBC, (uint64_t)B, (uint64_t)N, (uint64_t)T};
    uint64_t str[3] = {(uint64_t)fusedh::IBC * 2, (uint64_t)B * fusedh::IBC * 2, R * fusedh::IBC * 2};
    MCRN_TRY(fusedh::encode_tensor_map_h(&tIB, ib16_all ? ib16_all : x16_all, dims, str, box));
  }
  {
    uint64_t dims[4] = {(uint64_t)HS, (uint64_t)B, (uint64_t)N, (uint64_t)T};
    uint64_t str[3] = {(uint64_t)HS * 2, (uint64_t)B * HS * 2, R * HS * 2};
    MCRN_TRY(fusedh::encode_tensor_map_h(&tX, x16_all, dims, str, box));
  }
  {
    uint64_t dims[4] = {(uint64_t)O, (uint64_t)B, (uint64_t)N, (uint64_t)T};
    uint64_t str[3] = {(uint64_t)O * 2, (uint64_t)B * O * 2, R * O * 2};
    MCRN_TRY(fusedh::encode_tensor_map_h(&tV, v16_all, dims, str, box));
  }
  {
    uint64_t dims[4] = {(uint64_t)HS, (uint64_t)B, (uint64_t)N, (uint64_t)T * nks};
    uint64_t str[3] = {(uint64_t)HS * 2, (uint64_t)B * HS * 2, R * HS * 2};
    MCRN_TRY(fusedh::encode_tensor_map_h(&tQ, q16_all, dims, str, box));
  }
  WParams p;
  p.N = N; p.B = B; p.T = T; p.KS = KS; p.nhalf = nhalf; p.O = O; p.dw = dw; p.gs = gs;
  auto kern = agcn_dw_h_kernel<HS>;
  static bool attr_set = false;
  if (!attr_set) {
    MCRN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    attr_set = true;
  }
  const int tiles = (1 + KS + (ib16_all ? 1 : 0)) * nhalf;
  int G = 148 / tiles;
  if (G < 1) G = 1;
  if (G > T * B) G = T * B;
  MCRN_LAUNCH(kern, dim3(tiles, G), WTHREADS, C::SMEM, st, tX, tV, tQ, tIB, p);
  return MCRN_OK;
}

}  // namespace fusedwh
}  // namespace mcrn
