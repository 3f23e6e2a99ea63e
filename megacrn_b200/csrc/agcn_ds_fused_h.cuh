// Fused support-gradient kernel, fp16-operand version:  dS_k += sum_{t,b} (dV_{t,b} W_k^T) X_{t,b}^T  for ONE AGCN type over
// ALL time steps and batch elements in one launch.  Same structure as agcn_ds_fused.cuh (CTA = (128-node tile n, support k,
// group g) looping over its (t, b) units; dXP only in tensor memory), with the operands of the fp16 backward:
//     dV16 [T][R][O]     scaled fp16 gradient copies of every step (written by k_bwd_glue_h / EpiBUH)   A of MMA1
//     W16n [KS+2][HS][O] fp16 folded weights                                                            B of MMA1
//     X16  [T][R][HS]    fp16 state copies of every step (written by the fused forward)                 B of MMA2
// MMA1 dXP[128 x HS] = dV16 tile * W_k^T (kind::f16, fp32 in TMEM) -> fp16 pairs packed in place -> MMA2 acc[128 x N] +=
// dXP * X^T (A from TMEM).  dV16 carries the loss scale gs[0]; the accumulator is multiplied by gs[1] = 1/scale before
// the atomic add.  Half the operand bytes and twice the MMA rate of the TF32 version.  Requires N <= 256.
#pragma once

#include "agcn_bwd_fused_h.cuh"
#include "agcn_ds_fused.cuh"

namespace mcrn {
namespace fuseddh {

using namespace tc;
using fused::mbar_arrive;
using fused::mbar_wait_b;
using fused::tmem_wait_st;
using fusedd::D_ITEM_P;
using fusedd::D_ITEM_TS;
using fusedh::BKH;
using fusedh::make_idesc_f16;
using fusedh::pack_h2;
using fusedh::tcgen05_mma_f16;
using fusedh::tcgen05_mma_f16_ts;
using fusedh::tmem_st_32x32b_x16;

constexpr int DHTHREADS = 320;

struct DHParams {
  int N, B, T, O;
  int npad;              // ceil16(N)
  float* dS;             // [KS][N][ldS]
  int ldS;
  const float* gs;       // device {scale, 1 / scale}
};

template <int HS>
struct CfgDH {
  static_assert(HS == 64 || HS == 128, "hidden width: 64 or 128");
  static constexpr uint32_t A_SLOT = BM * 128;                    // [128 rows][64 halves]
  static constexpr uint32_t B_SLOT = 256 * 128;                   // up to 256 rows x 128 B (X rows; W needs HS rows)
  static constexpr uint32_t STAGE = A_SLOT + B_SLOT;              // 48 KB
  static constexpr int NST = 4;
  static constexpr size_t SMEM = (size_t)NST * STAGE + 1024;
  static constexpr uint32_t TM_ACC = 0, TM_Q0 = 256, TM_Q1 = 256 + HS;
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr int KB2 = HS / BKH;
};

template <int HS>
__global__ void __launch_bounds__(DHTHREADS, 1)
agcn_ds_h_kernel(const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmX, DHParams p) {
  using C = CfgDH<HS>;
  constexpr int NST = C::NST;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[NST];
  __shared__ __align__(8) uint64_t empty_bar[NST];
  __shared__ __align__(8) uint64_t q_full_bar[2];
  __shared__ __align__(8) uint64_t q_ready_bar[2];
  __shared__ __align__(8) uint64_t acc_full_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BM, k = blockIdx.y, grp = blockIdx.z, G = gridDim.z;
  const int U = p.T * p.B;
  const int nu = (U - grp + G - 1) / G;              // units grp, grp + G, ...
  const int kb1 = p.O / BKH;
  if (nu <= 0) return;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&q_full_bar[0]), 1);
    mbar_init(smem_u32(&q_full_bar[1]), 1);
    mbar_init(smem_u32(&q_ready_bar[0]), 4);
    mbar_init(smem_u32(&q_ready_bar[1]), 4);
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

  if (warp == 0) {
    if (lane == 0) {                                     // ===== TMA producer =====
      int it = 0;
      fusedd::for_each_item_d<C::KB2>(nu, kb1, [&](int type, int i, int j) {
        const int s = it % NST;
        if (it >= NST) mbar_wait_b(smem_u32(&empty_bar[s]), (((uint32_t)(it / NST)) & 1u) ^ 1u);
        const uint32_t fb = smem_u32(&full_bar[s]);
        const uint32_t a_dst = smem_base + (uint32_t)s * C::STAGE, b_dst = a_dst + C::A_SLOT;
        const int u = grp + i * G, t = u / p.B, b = u - t * p.B;
        if (type == D_ITEM_P) {
          mbar_expect_tx(fb, C::A_SLOT + (uint32_t)HS * 128);
          tma_load_4d(a_dst, &tmV, fb, j * BKH, b, n0, t);                       // dV16_t[n0.., b, 64 j..]
          tma_load_4d(b_dst, &tmW, fb, j * BKH, 0, 1 + k, 0);                    // W_k[0..HS][64 j..]
        } else {
          mbar_expect_tx(fb, (uint32_t)p.npad * 128);
          tma_load_4d(b_dst, &tmX, fb, j * BKH, b, 0, t);                        // X16_t[0..npad, b, 64 j..]
        }
        ++it;
      });
    }
  } else if (warp == 1) {
    if (lane == 0) {                                     // ===== MMA issuer =====
      constexpr uint32_t idesc1 = make_idesc_f16<HS>();
      const uint32_t idesc2 = (1u << 4) | ((uint32_t)(p.npad >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int it = 0;
      bool acc_on = false;
      fusedd::for_each_item_d<C::KB2>(nu, kb1, [&](int type, int i, int j) {
        const int s = it % NST;
        mbar_wait_b(smem_u32(&full_bar[s]), ((uint32_t)(it / NST)) & 1u);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_base + (uint32_t)s * C::STAGE, b_addr = a_addr + C::A_SLOT;
        const uint32_t qbuf = tmem_base + ((i & 1) ? C::TM_Q1 : C::TM_Q0);
        if (type == D_ITEM_P) {
#pragma unroll
          for (int kk = 0; kk < BKH / 16; ++kk) {
            const uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024, 2);
            const uint64_t bd = make_smem_desc(b_addr + kk * 32, 16, 1024, 2);
            tcgen05_mma_f16(qbuf, ad, bd, idesc1, (j > 0 || kk > 0) ? 1u : 0u);
          }
          tcgen05_commit(smem_u32(&empty_bar[s]));
          if (j == kb1 - 1) tcgen05_commit(smem_u32(&q_full_bar[i & 1]));
        } else {
          if (j == 0) {
            mbar_wait_b(smem_u32(&q_ready_bar[i & 1]), ((uint32_t)(i >> 1)) & 1u);
            tcgen05_fence_after();
          }
#pragma unroll
          for (int kk = 0; kk < BKH / 16; ++kk) {
            const uint64_t bd = make_smem_desc(b_addr + kk * 32, 16, 1024, 2);
            tcgen05_mma_f16_ts(tmem_base + C::TM_ACC, qbuf + (uint32_t)(j * (BKH / 2) + kk * 8), bd, idesc2, (acc_on || kk > 0) ? 1u : 0u);
          }
          acc_on = true;
          tcgen05_commit(smem_u32(&empty_bar[s]));
        }
        ++it;
      });
      tcgen05_commit(smem_u32(&acc_full_bar));
    }
  } else {                                               // ===== packing (warps 2..5) + epilogue (warps 2..9) =====
    const int quarter = warp & 3;
    const int ew = warp - 2, half_id = ew >> 2;
    const int cq = (lane & 7) * 4, r0 = lane >> 3;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int node0 = n0 + quarter * 32;
    if (half_id == 0) {
      // dXP: fp32 accumulator -> fp16 pairs packed in place (one warp per lane quarter: chunk c packs into columns
      // [16c, 16c+16), which this warp has already read)
      for (int i = 0; i < nu; ++i) {
        mbar_wait_b(smem_u32(&q_full_bar[i & 1]), ((uint32_t)(i >> 1)) & 1u);
        tcgen05_fence_after();
        const uint32_t qbuf = tmem_base + ((i & 1) ? C::TM_Q1 : C::TM_Q0) + lane_off;
#pragma unroll 1
        for (int c = 0; c < HS / 32; ++c) {
          float v[32];
          tmem_ld_32x32b_x32(qbuf + (uint32_t)(c * 32), v);
          uint32_t u[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) u[e] = pack_h2(v[2 * e], v[2 * e + 1]);
          tmem_st_32x32b_x16(qbuf + (uint32_t)(c * 16), u);
        }
        tmem_wait_st();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&q_ready_bar[i & 1]));
      }
    }
    mbar_wait_b(smem_u32(&acc_full_bar), 0);
    tcgen05_fence_after();
    if (node0 < p.N) {
      const float inv_gs = __ldg(p.gs + 1);
      float* scr = reinterpret_cast<float*>(smem_al) + ew * (32 * 36);      // the ring is idle now
      float* dst_k = p.dS + (int64_t)k * p.N * p.ldS;
#pragma unroll 1
      for (int c = half_id; c * 32 < p.npad; c += 2) {
        float v[32];
        tmem_ld_32x32b_x32(tmem_base + C::TM_ACC + lane_off + (uint32_t)(c * 32), v);
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 32; e += 4)
          *reinterpret_cast<float4*>(&scr[lane * 36 + e]) = make_float4(v[e] * inv_gs, v[e + 1] * inv_gs, v[e + 2] * inv_gs, v[e + 3] * inv_gs);
        __syncwarp();
        const int col = c * 32 + cq;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int rr = r0 + 4 * e, node = node0 + rr;
          if (node < p.N && col < p.ldS) {
            const float4 t4 = *reinterpret_cast<const float4*>(&scr[rr * 36 + cq]);
            atomicAdd(reinterpret_cast<float4*>(dst_k + (int64_t)node * p.ldS + col), t4);
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

// dV16_all: [T][R][O] scaled fp16.  w16n: [KS+2][HS][O] fp16.  x16_all: [T][R][HS] fp16.  dS: [KS][N][ldS] fp32, atomics.
template <int HS>
int launch_agcn_ds_h(int N, int B, int T, int KS, int ldS, int O, const __half* dV16_all, const __half* w16n, const __half* x16_all,
                     const float* gs, float* dS, cudaStream_t st) {
  using C = CfgDH<HS>;
  const int64_t R = (int64_t)N * B;
  const int npad = (N + 15) / 16 * 16;
  CUtensorMap tV, tW, tX;
  {
    uint64_t dims[4] = {(uint64_t)O, (uint64_t)B, (uint64_t)N, (uint64_t)T};
    uint64_t str[3] = {(uint64_t)O * 2, (uint64_t)B * O * 2, (uint64_t)R * O * 2};
    uint32_t box[4] = {BKH, 1, BM, 1};
    MCRN_TRY(fusedh::encode_tensor_map_h(&tV, dV16_all, dims, str, box));
  }
  {
    uint64_t dims[4] = {(uint64_t)O, (uint64_t)HS, (uint64_t)(KS + 2), 1};
    uint64_t str[3] = {(uint64_t)O * 2, (uint64_t)HS * O * 2, (uint64_t)(KS + 2) * HS * O * 2};
    uint32_t box[4] = {BKH, (uint32_t)HS, 1, 1};
    MCRN_TRY(fusedh::encode_tensor_map_h(&tW, w16n, dims, str, box));
  }
  {
    uint64_t dims[4] = {(uint64_t)HS, (uint64_t)B, (uint64_t)N, (uint64_t)T};
    uint64_t str[3] = {(uint64_t)HS * 2, (uint64_t)B * HS * 2, (uint64_t)R * HS * 2};
    uint32_t box[4] = {BKH, 1, (uint32_t)npad, 1};
    MCRN_TRY(fusedh::encode_tensor_map_h(&tX, x16_all, dims, str, box));
  }
  DHParams p;
  p.N = N; p.B = B; p.T = T; p.O = O; p.npad = npad; p.dS = dS; p.ldS = ldS; p.gs = gs;
  auto kern = agcn_ds_h_kernel<HS>;
  static bool attr_set = false;
  if (!attr_set) {
    MCRN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    attr_set = true;
  }
  const int tiles = ceil_div(N, BM);
  int G = 148 / (tiles * KS);
  if (G < 1) G = 1;
  if (G > T * B) G = T * B;
  dim3 grid(tiles, KS, G);
  MCRN_LAUNCH(kern, grid, DHTHREADS, C::SMEM, st, tV, tW, tX, p);
  return MCRN_OK;
}

}  // namespace fuseddh
}  // namespace mcrn
