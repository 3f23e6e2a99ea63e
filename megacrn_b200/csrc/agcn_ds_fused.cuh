// Fused support-gradient kernel for sm_100a:   dS_k += sum_{t,b} (dV_{t,b} W_k^T) X_{t,b}^T      (tests/kernel_spec.py:d_supports)
// for ONE AGCN type (encoder/decoder x gate/update) over ALL time steps and batch elements in one launch.  The
// [N x HS] intermediate dXP_k = dV W_k^T of every (t, b) -- 27 MB per AGCN call at C2 when it goes through HBM/L2 --
// lives only in tensor memory:
//
//   CTA (128-node tile n, support k, group g) loops over its (t, b) units:
//     MMA1  dXP[128 x HS]  = dV_t[tile rows, b, :] * W_k^T        A = dV rows (TMA, K-major), B = folded weights [c][o] (K-major)
//     round dXP -> TF32 in place in TMEM (two buffers, ping-pong over units)
//     MMA2  acc[128 x N]  += dXP * X_t[:, b, :]^T                 A = dXP from TMEM, B = state rows of b (TMA, K-major)
//   epilogue (once): atomically add the accumulator tile to dS_k (one red.global.add.v4.f32 per 4 columns).
//
// Requires N <= 256 (the dS tile occupies ceil16(N) TMEM columns).  Warp roles as in agcn_bwd_fused.cuh.
#pragma once

#include "agcn_bwd_fused.cuh"

namespace mcrn {
namespace fusedd {

using namespace tc;
using fused::mbar_arrive;
using fused::mbar_wait_b;
using fused::tcgen05_mma_tf32_ts;
using fused::tmem_st_32x32b_x32;
using fused::tmem_wait_st;

constexpr int DTHREADS = 320;

struct DParams {
  int N, B, T, O;        // nodes, batch, steps, width of dV (HS or 2 HS)
  int npad;              // ceil16(N): MMA2 N / rows of the X box
  int k0;                // weight segment of support 0 (= 1)
  float* dS;             // [KS][N][ldS]
  int ldS;
};

template <int HS>
struct CfgD {
  static_assert(HS == 64 || HS == 128, "hidden width: 64 or 128");
  static constexpr uint32_t A_SLOT = BM * BK * 4;                 // 16 KB
  static constexpr uint32_t B_SLOT = 256 * BK * 4;                // up to 256 rows x 128 B (X rows; W needs HS rows)
  static constexpr uint32_t STAGE = A_SLOT + B_SLOT;              // 48 KB
  static constexpr int NST = 4;
  static constexpr size_t SMEM = (size_t)NST * STAGE + 1024;
  static constexpr uint32_t TM_ACC = 0, TM_Q0 = 256, TM_Q1 = 256 + HS;
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr int KB2 = HS / BK;
};

enum : int { D_ITEM_P = 0, D_ITEM_TS = 1 };

// ring items in issue order over the CTA's local units i = 0..nu-1
template <int KB2, class F>
__device__ __forceinline__ void for_each_item_d(int nu, int kb1, F&& f) {
  if (nu <= 0) return;
  for (int j = 0; j < kb1; ++j) f(D_ITEM_P, 0, j);
  if (nu > 1)
    for (int j = 0; j < kb1; ++j) f(D_ITEM_P, 1, j);
  for (int i = 0; i < nu; ++i) {
    for (int j = 0; j < KB2; ++j) f(D_ITEM_TS, i, j);
    if (i + 2 < nu)
      for (int j = 0; j < kb1; ++j) f(D_ITEM_P, i + 2, j);
  }
}

template <int HS>
__global__ void __launch_bounds__(DTHREADS, 1)
agcn_ds_kernel(const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmX, DParams p) {
  using C = CfgD<HS>;
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
  const int nu = (U - grp + G - 1) / G;              // units grp, grp + G, ...   (<= 0: nothing to do)
  const int kb1 = p.O / BK;
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
    mbar_init(smem_u32(&q_ready_bar[0]), 8);
    mbar_init(smem_u32(&q_ready_bar[1]), 8);
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
      for_each_item_d<C::KB2>(nu, kb1, [&](int type, int i, int j) {
        const int s = it % NST;
        if (it >= NST) mbar_wait_b(smem_u32(&empty_bar[s]), (((uint32_t)(it / NST)) & 1u) ^ 1u);
        const uint32_t fb = smem_u32(&full_bar[s]);
        const uint32_t a_dst = smem_base + (uint32_t)s * C::STAGE, b_dst = a_dst + C::A_SLOT;
        const int u = grp + i * G, t = u / p.B, b = u - t * p.B;
        if (type == D_ITEM_P) {
          mbar_expect_tx(fb, C::A_SLOT + (uint32_t)HS * BK * 4);
          tma_load_4d(a_dst, &tmV, fb, j * BK, b, n0, t);                        // dV_t[n0.., b, 32 j..]
          tma_load_4d(b_dst, &tmW, fb, j * BK, 0, p.k0 + k, 0);                  // W_k[0..HS][32 j..]
        } else {
          mbar_expect_tx(fb, (uint32_t)p.npad * BK * 4);
          tma_load_4d(b_dst, &tmX, fb, j * BK, b, 0, t);                         // X_t[0..npad, b, 32 j..]
        }
        ++it;
      });
    }
  } else if (warp == 1) {
    if (lane == 0) {                                     // ===== MMA issuer =====
      constexpr uint32_t idesc1 = make_idesc<true, true, HS>();
      const uint32_t idesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.npad >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int it = 0;
      bool acc_on = false;
      for_each_item_d<C::KB2>(nu, kb1, [&](int type, int i, int j) {
        const int s = it % NST;
        mbar_wait_b(smem_u32(&full_bar[s]), ((uint32_t)(it / NST)) & 1u);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_base + (uint32_t)s * C::STAGE, b_addr = a_addr + C::A_SLOT;
        const uint32_t qbuf = tmem_base + ((i & 1) ? C::TM_Q1 : C::TM_Q0);
        if (type == D_ITEM_P) {
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024, 2);
            const uint64_t bd = make_smem_desc(b_addr + kk * 32, 16, 1024, 2);
            tcgen05_mma_tf32(qbuf, ad, bd, idesc1, (j > 0 || kk > 0) ? 1u : 0u);
          }
          tcgen05_commit(smem_u32(&empty_bar[s]));
          if (j == kb1 - 1) tcgen05_commit(smem_u32(&q_full_bar[i & 1]));
        } else {
          if (j == 0) {
            mbar_wait_b(smem_u32(&q_ready_bar[i & 1]), ((uint32_t)(i >> 1)) & 1u);
            tcgen05_fence_after();
          }
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t bd = make_smem_desc(b_addr + kk * 32, 16, 1024, 2);
            tcgen05_mma_tf32_ts(tmem_base + C::TM_ACC, qbuf + (uint32_t)(j * BK + kk * 8), bd, idesc2, (acc_on || kk > 0) ? 1u : 0u);
          }
          acc_on = true;
          tcgen05_commit(smem_u32(&empty_bar[s]));
        }
        ++it;
      });
      tcgen05_commit(smem_u32(&acc_full_bar));
    }
  } else {                                               // ===== rounding + epilogue warps =====
    const int quarter = warp & 3;
    const int ew = warp - 2, half_id = ew >> 2;
    const int cq = (lane & 7) * 4, r0 = lane >> 3;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int node0 = n0 + quarter * 32;
    for (int i = 0; i < nu; ++i) {
      mbar_wait_b(smem_u32(&q_full_bar[i & 1]), ((uint32_t)(i >> 1)) & 1u);
      tcgen05_fence_after();
      const uint32_t qbuf = tmem_base + ((i & 1) ? C::TM_Q1 : C::TM_Q0) + lane_off;
#pragma unroll 1
      for (int c = half_id; c < HS / 32; c += 2) {
        float v[32];
        tmem_ld_32x32b_x32(qbuf + (uint32_t)(c * 32), v);
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = tf32_rn(v[e]);
        tmem_st_32x32b_x32(qbuf + (uint32_t)(c * 32), v);
      }
      tmem_wait_st();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&q_ready_bar[i & 1]));
    }
    mbar_wait_b(smem_u32(&acc_full_bar), 0);
    tcgen05_fence_after();
    if (node0 < p.N) {
      float* scr = reinterpret_cast<float*>(smem_al) + ew * (32 * 36);      // the ring is idle now
      float* dst_k = p.dS + (int64_t)k * p.N * p.ldS;
#pragma unroll 1
      for (int c = half_id; c * 32 < p.npad; c += 2) {
        float v[32];
        tmem_ld_32x32b_x32(tmem_base + C::TM_ACC + lane_off + (uint32_t)(c * 32), v);
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 32; e += 4) *reinterpret_cast<float4*>(&scr[lane * 36 + e]) = make_float4(v[e], v[e + 1], v[e + 2], v[e + 3]);
        __syncwarp();
        const int col = c * 32 + cq;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int rr = r0 + 4 * e, node = node0 + rr;
          if (node < p.N && col < p.ldS) {            // columns N..ldS-1 of the accumulator are zero (OOB rows of X)
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

static inline bool ds_fused_eligible(int N, int HS) { return (HS == 64 || HS == 128) && N <= 256; }

// dV_all: [T][R][O] (TF32-rounded).  wall: folded weights (hi part) [KS+2][HS][O].  xp0: XP block 0 of step 0, steps
// xp_step floats apart ([R][HS] each).  dS: [KS][N][ldS] accumulated atomically.
template <int HS>
int launch_agcn_ds(int N, int B, int T, int KS, int ldS, int O, const float* dV_all, const float* wall, const float* xp0,
                   int64_t xp_step, float* dS, cudaStream_t st) {
  using C = CfgD<HS>;
  const int64_t R = (int64_t)N * B;
  const int npad = (N + 15) / 16 * 16;
  CUtensorMap tV, tW, tX;
  {
    uint64_t dims[4] = {(uint64_t)O, (uint64_t)B, (uint64_t)N, (uint64_t)T};
    uint64_t str[3] = {(uint64_t)O * 4, (uint64_t)B * O * 4, (uint64_t)R * O * 4};
    uint32_t box[4] = {32, 1, BM, 1};
    MCRN_TRY(encode_tensor_map(&tV, dV_all, dims, str, box, false));
  }
  {
    uint64_t dims[4] = {(uint64_t)O, (uint64_t)HS, (uint64_t)(KS + 2), 1};
    uint64_t str[3] = {(uint64_t)O * 4, (uint64_t)HS * O * 4, (uint64_t)(KS + 2) * HS * O * 4};
    uint32_t box[4] = {32, (uint32_t)HS, 1, 1};
    MCRN_TRY(encode_tensor_map(&tW, wall, dims, str, box, false));
  }
  {
    uint64_t dims[4] = {(uint64_t)HS, (uint64_t)B, (uint64_t)N, (uint64_t)T};
    uint64_t str[3] = {(uint64_t)HS * 4, (uint64_t)B * HS * 4, (uint64_t)xp_step * 4};
    uint32_t box[4] = {32, 1, (uint32_t)npad, 1};
    MCRN_TRY(encode_tensor_map(&tX, xp0, dims, str, box, false));
  }
  DParams p;
  p.N = N; p.B = B; p.T = T; p.O = O; p.npad = npad; p.k0 = 1; p.dS = dS; p.ldS = ldS;
  auto kern = agcn_ds_kernel<HS>;
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
  MCRN_LAUNCH(kern, grid, DTHREADS, C::SMEM, st, tV, tW, tX, p);
  return MCRN_OK;
}

}  // namespace fusedd
}  // namespace mcrn
