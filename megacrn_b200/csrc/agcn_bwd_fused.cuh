// Fused AGCN backward (data path) for sm_100a: for ONE AGCN call, the gradient w.r.t. its state operand
//     dX = dV W_0^T + sum_k S_k^T (dV W_k^T)  =  dV W_0^T + sum_k (S_k^T dV) W_k^T
// is evaluated in the second form, which has the same chained shape as the forward kernel (agcn_fused.cuh):
//
//   CTA (128-node tile m, batch element b), "supports" ks = (k, half) -- the 2H-wide dG of the gate AGCN is
//   processed as two H-wide halves so that every propagated block is 128 x HS:
//     MMA1  Q_ks[128 x HS] = S_k^T[tile rows, :] * dV[:, b, half]      A = transposed supports (TMA, K-major)
//                                                                       B = dV of batch element b (TMA, MN-major slabs)
//     round Q_ks -> TF32 in place in TMEM; store it (dW_k = sum X^T Q_k is one GEMM per AGCN at the end)
//     MMA2  acc[128 x HS] += Q_ks * W_k[:, half]^T   (A = Q_ks from TMEM; B = folded weights [c][o], K-major)
//                          + dV_tile * W_0^T          (A in smem)
//           accIB[128 x 16] = dV_tile * W_NB[0..16)^T (gradient of the input-channel block, first NB*Cin columns)
//     epilogue  update AGCN: the gate backward (dG, dh_part) -- EpiBU;  gate AGCN: dH_prev = acc + dh_part -- EpiBG
//
// tests/kernel_spec.py:cell_bwd is the algebra; model.cu:cell_backward_fused the orchestration.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = rounding + epilogue
// (two warps per TMEM lane quarter, splitting the 32-column chunks).
#pragma once

#include "agcn_fused.cuh"

namespace mcrn {
namespace fusedb {

using namespace tc;
using fused::mbar_arrive;
using fused::mbar_wait_b;
using fused::pow2_cols;
using fused::tcgen05_mma_tf32_ts;
using fused::tmem_st_32x32b_x32;
using fused::tmem_wait_st;

constexpr int BTHREADS = 320;
constexpr int IBW = 16;                          // columns of the input-block gradient that are computed

struct BParams {
  int N, B, KS, nhalf;   // nhalf = 1: dV is [R][HS] (update AGCN); 2: dV is [R][2 HS] (gate AGCN)
  float* qsave;          // [KS * nhalf][R][HS]: rounded Q blocks, block index ks = k * nhalf + half
  int64_t blk_stride;    // R * HS
  float* dib;            // [R][IBW] gradient of the input block (columns < IBW)
};

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int HS>
struct CfgB {
  static_assert(HS == 64 || HS == 128, "hidden width of the fused AGCN backward kernel: 64 or 128");
  static constexpr uint32_t A_SLOT = BM * BK * 4;                 // 16 KB
  static constexpr uint32_t B_SLOT = (uint32_t)HS * BK * 4;       // HS x 32 floats (slabs for MMA1, [HS rows][128 B] for MMA2)
  static constexpr uint32_t IB_SLOT = IBW * BK * 4;               // 2 KB: [16 rows][128 B]
  static constexpr uint32_t STAGE = A_SLOT + B_SLOT + IB_SLOT;
  static constexpr int NST = HS >= 128 ? 5 : 6;
  static constexpr uint32_t SCRATCH = 8 * 32 * 36 * 4;
  static constexpr size_t SMEM = (size_t)NST * STAGE + SCRATCH + 1024;
  static constexpr uint32_t TM_ACC = 0, TM_IB = HS, TM_Q0 = HS + 32, TM_Q1 = 2 * HS + 32;
  static constexpr uint32_t TMEM_COLS = pow2_cols(3 * HS + 32);
  static constexpr int KB2 = HS / BK;
};

enum : int { B_ITEM_P = 0, B_ITEM_SS = 1, B_ITEM_TS = 2 };

// ring items in issue order; ks = k * nhalf + half
template <int KB2, class F>
__device__ __forceinline__ void for_each_item_b(int nks, int nhalf, int kb1, F&& f) {
  for (int j = 0; j < kb1; ++j) f(B_ITEM_P, 0, j);
  if (nks > 1)
    for (int j = 0; j < kb1; ++j) f(B_ITEM_P, 1, j);
  for (int h = 0; h < nhalf; ++h)
    for (int j = 0; j < KB2; ++j) f(B_ITEM_SS, h, j);
  for (int ks = 0; ks < nks; ++ks) {
    for (int j = 0; j < KB2; ++j) f(B_ITEM_TS, ks, j);
    if (ks + 2 < nks)
      for (int j = 0; j < kb1; ++j) f(B_ITEM_P, ks + 2, j);
  }
}

// ---- epilogue functors (4 consecutive columns of one (node, b) row) -----------------------------
// Update-AGCN backward tail = the gate backward of the cell (tests/kernel_spec.py:cell_bwd):
//   dZH = acc ; dG[:, :H] = dZH*h*z(1-z) ; dh_part = dH'*r + dZH*z
// (dG[:, H:] = dH'*(h-hc)*r(1-r) and dHr = dH'*r do not depend on dZH: k_bwd_glue has already written them)
struct EpiBU {
  static constexpr int NP = 3;
  int H;
  const float *z, *h, *dHr;
  float *dG, *dh_part;
  __device__ __forceinline__ void load4(int row, int n0, float4 (&p)[NP]) const {
    const int64_t f = (int64_t)row * H + n0;
    p[0] = ldg4(z + f); p[1] = ldg4(h + f); p[2] = ldg4(dHr + f);
  }
  __device__ __forceinline__ void fin4(int row, int n0, const float4 (&p)[NP], const float (&acc)[4]) const {
    const float4 zz = p[0], hh = p[1], dr = p[2];
    const float d0 = acc[0], d1 = acc[1], d2 = acc[2], d3 = acc[3];
    st4(dG + (int64_t)row * 2 * H + n0, tf32_rn(d0 * hh.x * zz.x * (1.0f - zz.x)), tf32_rn(d1 * hh.y * zz.y * (1.0f - zz.y)),
        tf32_rn(d2 * hh.z * zz.z * (1.0f - zz.z)), tf32_rn(d3 * hh.w * zz.w * (1.0f - zz.w)));
    st4(dh_part + (int64_t)row * H + n0, dr.x + d0 * zz.x, dr.y + d1 * zz.y, dr.z + d2 * zz.z, dr.w + d3 * zz.w);
  }
};
// Gate-AGCN backward tail: dH_prev = acc + dh_part
struct EpiBG {
  static constexpr int NP = 1;
  int H;
  const float* dh_part;
  float* dH_out;
  __device__ __forceinline__ void load4(int row, int n0, float4 (&p)[NP]) const { p[0] = ldg4(dh_part + (int64_t)row * H + n0); }
  __device__ __forceinline__ void fin4(int row, int n0, const float4 (&p)[NP], const float (&acc)[4]) const {
    st4(dH_out + (int64_t)row * H + n0, acc[0] + p[0].x, acc[1] + p[0].y, acc[2] + p[0].z, acc[3] + p[0].w);
  }
};

template <int HS, class Epi>
__global__ void __launch_bounds__(BTHREADS, 1)
agcn_bwd_kernel(const __grid_constant__ CUtensorMap tmST, const __grid_constant__ CUtensorMap tmVB,
                const __grid_constant__ CUtensorMap tmVA, const __grid_constant__ CUtensorMap tmW,
                const __grid_constant__ CUtensorMap tmWib, BParams p, Epi epi) {
  using C = CfgB<HS>;
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
  const int m0 = blockIdx.x * BM, b = blockIdx.y;
  const int kb1 = (p.N + BK - 1) / BK;
  const int nks = p.KS * p.nhalf;
  const int NBLK = p.KS + 1;                          // weight segment of the input block

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmST) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmVB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmVA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmWib) : "memory");
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
      for_each_item_b<C::KB2>(nks, p.nhalf, kb1, [&](int type, int ks, int j) {
        const int s = it % NST;
        if (it >= NST) mbar_wait_b(smem_u32(&empty_bar[s]), (((uint32_t)(it / NST)) & 1u) ^ 1u);
        const uint32_t fb = smem_u32(&full_bar[s]);
        const uint32_t a_dst = smem_base + (uint32_t)s * C::STAGE, b_dst = a_dst + C::A_SLOT, ib_dst = b_dst + C::B_SLOT;
        if (type == B_ITEM_P) {
          const int k = ks / p.nhalf, half = ks - k * p.nhalf;
          mbar_expect_tx(fb, C::A_SLOT + C::B_SLOT);
          tma_load_4d(a_dst, &tmST, fb, j * BK, m0, k, 0);                         // S_k^T[m0.., 32 j..]
#pragma unroll
          for (int q = 0; q < HS / 32; ++q)                                         // dV[32 j.., b, half*HS + 32 q..]
            tma_load_4d(b_dst + q * SLAB_BYTES, &tmVB, fb, half * HS + 32 * q, b, j * BK, 0);
        } else if (type == B_ITEM_SS) {
          const int half = ks;
          mbar_expect_tx(fb, C::A_SLOT + C::B_SLOT + C::IB_SLOT);
          tma_load_4d(a_dst, &tmVA, fb, half * HS + j * BK, b, m0, 0);             // dV tile, columns (o) 32 j..
          tma_load_4d(b_dst, &tmW, fb, half * HS + j * BK, 0, 0, 0);               // W_0[0..HS][o-block]
          tma_load_4d(ib_dst, &tmWib, fb, half * HS + j * BK, 0, NBLK, 0);         // W_NB[0..16][o-block]
        } else {
          const int k = ks / p.nhalf, half = ks - k * p.nhalf;
          mbar_expect_tx(fb, C::B_SLOT);
          tma_load_4d(b_dst, &tmW, fb, half * HS + j * BK, 0, 1 + k, 0);           // W_{1+k}[0..HS][o-block]
        }
        ++it;
      });
    }
  } else if (warp == 1) {
    if (lane == 0) {                                     // ===== MMA issuer =====
      constexpr uint32_t idesc1 = make_idesc<true, false, HS>();     // A K-major, B MN-major (dV slabs)
      constexpr uint32_t idesc2 = make_idesc<true, true, HS>();      // B K-major (weights [c][o])
      constexpr uint32_t idesc3 = make_idesc<true, true, IBW>();
      int it = 0;
      bool acc_on = false, ib_on = false;
      for_each_item_b<C::KB2>(nks, p.nhalf, kb1, [&](int type, int ks, int j) {
        const int s = it % NST;
        mbar_wait_b(smem_u32(&full_bar[s]), ((uint32_t)(it / NST)) & 1u);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_base + (uint32_t)s * C::STAGE, b_addr = a_addr + C::A_SLOT, ib_addr = b_addr + C::B_SLOT;
        const uint32_t qbuf = tmem_base + ((ks & 1) ? C::TM_Q1 : C::TM_Q0);
        if (type == B_ITEM_P) {
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024, 2);
            const uint64_t bd = make_smem_desc(b_addr + kk * 1024, SLAB_BYTES, 512, 1);
            tcgen05_mma_tf32(qbuf, ad, bd, idesc1, (j > 0 || kk > 0) ? 1u : 0u);
          }
          tcgen05_commit(smem_u32(&empty_bar[s]));
          if (j == kb1 - 1) tcgen05_commit(smem_u32(&q_full_bar[ks & 1]));
        } else if (type == B_ITEM_SS) {
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024, 2);
            const uint64_t bd = make_smem_desc(b_addr + kk * 32, 16, 1024, 2);
            const uint64_t id = make_smem_desc(ib_addr + kk * 32, 16, 1024, 2);
            tcgen05_mma_tf32(tmem_base + C::TM_ACC, ad, bd, idesc2, (acc_on || kk > 0) ? 1u : 0u);
            tcgen05_mma_tf32(tmem_base + C::TM_IB, ad, id, idesc3, (ib_on || kk > 0) ? 1u : 0u);
          }
          acc_on = true; ib_on = true;
          tcgen05_commit(smem_u32(&empty_bar[s]));
        } else {
          if (j == 0) {
            mbar_wait_b(smem_u32(&q_ready_bar[ks & 1]), ((uint32_t)(ks >> 1)) & 1u);
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
    const int node0 = m0 + quarter * 32;
    {
      float* scr = reinterpret_cast<float*>(smem_al + (size_t)NST * C::STAGE) + ew * (32 * 36);
      for (int ks = 0; ks < nks; ++ks) {
        mbar_wait_b(smem_u32(&q_full_bar[ks & 1]), ((uint32_t)(ks >> 1)) & 1u);
        tcgen05_fence_after();
        const uint32_t qbuf = tmem_base + ((ks & 1) ? C::TM_Q1 : C::TM_Q0) + lane_off;
#pragma unroll 1
        for (int c = half_id; c < HS / 32; c += 2) {
          float v[32];
          tmem_ld_32x32b_x32(qbuf + (uint32_t)(c * 32), v);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = tf32_rn(v[i]);
          tmem_st_32x32b_x32(qbuf + (uint32_t)(c * 32), v);
          if (node0 < p.N) {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(&scr[lane * 36 + i]) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            __syncwarp();
            float* dst = p.qsave + (int64_t)ks * p.blk_stride + c * 32 + cq;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = r0 + 4 * i, node = node0 + rr;
              if (node < p.N)
                *reinterpret_cast<float4*>(dst + ((int64_t)node * p.B + b) * HS) = *reinterpret_cast<const float4*>(&scr[rr * 36 + cq]);
            }
          }
        }
        tmem_wait_st();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&q_ready_bar[ks & 1]));
      }
    }
    mbar_wait_b(smem_u32(&acc_full_bar), 0);
    tcgen05_fence_after();
    if (node0 < p.N) {
      float* scr = reinterpret_cast<float*>(smem_al) + ew * (32 * 36);      // the ring is idle now
#pragma unroll 1
      for (int c = half_id; c < HS / 32; c += 2) {
        float v[32];
        tmem_ld_32x32b_x32(tmem_base + C::TM_ACC + lane_off + (uint32_t)(c * 32), v);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(&scr[lane * 36 + i]) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        __syncwarp();
        const int col = c * 32 + cq;
        constexpr int RB = Epi::NP <= 2 ? 8 : 4;
#pragma unroll
        for (int b0 = 0; b0 < 8; b0 += RB) {
          float4 pre[RB][Epi::NP];
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const int node = node0 + r0 + 4 * (b0 + i);
            if (node < p.N) epi.load4(node * p.B + b, col, pre[i]);
          }
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const int rr = r0 + 4 * (b0 + i), node = node0 + rr;
            if (node < p.N) {
              const float4 t = *reinterpret_cast<const float4*>(&scr[rr * 36 + cq]);
              const float a4[4] = {t.x, t.y, t.z, t.w};
              epi.fin4(node * p.B + b, col, pre[i], a4);
            }
          }
        }
      }
      if (half_id == 1) {                                // input-block gradient: 16 columns, one row per thread
        float v[16];
        tmem_ld_32x32b_x16(tmem_base + C::TM_IB + lane_off, v);
        const int node = node0 + lane;
        if (node < p.N) {
          float* dst = p.dib + ((int64_t)node * p.B + b) * IBW;
#pragma unroll
          for (int i = 0; i < 16; i += 4) st4(dst + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
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

// St[k][m][n] = S[k][n][m]  (the TF32-rounded supports, transposed once per backward)
__global__ void k_transpose_supports(const float* __restrict__ S, float* __restrict__ St, int n, int ld) {
  __shared__ float tile[32][33];
  const int k = blockIdx.z;
  const float* s = S + (int64_t)k * n * ld;
  float* t = St + (int64_t)k * n * ld;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = by + i, c = bx + threadIdx.x;
    tile[i][threadIdx.x] = (r < n && c < n) ? s[(int64_t)r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = bx + i, c = by + threadIdx.x;          // output row = input column
    if (r < n && c < ld) t[(int64_t)r * ld + c] = (c < n) ? tile[threadIdx.x][i] : 0.f;
  }
}

// Step glue of the fused backward (everything elementwise between two cells, one pass):
//   dH   = [dH_in] + d_out_t . wp                       projection backward (model/MegaCRN.py:186); dwp, dbp accumulate
//   dU   = dH * (1-r) * (1-hc^2)                        -> dV of the update AGCN (TF32-rounded)
//   dG_r = dH * (h-hc) * r(1-r)                         -> columns [D, 2D) of dG (TF32-rounded)
//   dHr  = dH * r                                       -> the part of dh_part that does not depend on dZH
// One block = 32 rows; a thread owns 4 consecutive columns.  D % 4 == 0, D <= 1024.
__global__ void __launch_bounds__(256) k_bwd_glue(const float* __restrict__ dOut, const float* __restrict__ dxin, int dxin_stride,
                                                  const float* __restrict__ h_t, const float* __restrict__ wp,
                                                  const float* __restrict__ dH, int dh_init, const float* __restrict__ r,
                                                  const float* __restrict__ hc, const float* __restrict__ hx,
                                                  float* __restrict__ dU, float* __restrict__ dG, float* __restrict__ dHr,
                                                  float* __restrict__ dwp, float* __restrict__ dbp, int B, int T, int N,
                                                  int D, int Cout, int t) {
  extern __shared__ float sh[];                    // [32][Cout] d_out rows, then [Cout][D] dwp partials
  float* sh_do = sh;
  float* sh_w = sh + 32 * Cout;
  const int64_t R = (int64_t)N * B, r0 = (int64_t)blockIdx.x * 32;
  const bool proj = (dOut != nullptr) || (dxin != nullptr);
  if (proj) {
    for (int i = threadIdx.x; i < 32 * Cout; i += blockDim.x) {
      const int64_t row = r0 + i / Cout;
      const int co = i % Cout;
      float v = 0.f;
      if (row < R) {
        const int n = (int)(row / B), b = (int)(row % B);
        if (dOut) v = dOut[(((int64_t)b * T + t) * N + n) * Cout + co];
        if (dxin) v += dxin[row * dxin_stride + co];
      }
      sh_do[i] = v;
    }
    for (int i = threadIdx.x; i < Cout * D; i += blockDim.x) sh_w[i] = 0.f;
    __syncthreads();
  }
  const int Q = D >> 2;
  for (int e = threadIdx.x; e < 32 * Q; e += blockDim.x) {
    const int i = e / Q, q4 = (e - i * Q) * 4;
    const int64_t row = r0 + i;
    if (row >= R) continue;
    const int64_t o = row * D + q4;
    float4 v = dh_init ? make_float4(0.f, 0.f, 0.f, 0.f) : ldg4(dH + o);
    if (proj) {
      const float4 hv = ldg4(h_t + o);
      for (int co = 0; co < Cout; ++co) {
        const float d = sh_do[i * Cout + co];
        const float4 w4 = ldg4(wp + (int64_t)co * D + q4);
        v.x = fmaf(d, w4.x, v.x); v.y = fmaf(d, w4.y, v.y); v.z = fmaf(d, w4.z, v.z); v.w = fmaf(d, w4.w, v.w);
        float* sw = sh_w + co * D + q4;
        atomicAdd(sw, d * hv.x); atomicAdd(sw + 1, d * hv.y); atomicAdd(sw + 2, d * hv.z); atomicAdd(sw + 3, d * hv.w);
      }
    }
    const float4 rr = ldg4(r + o), cc = ldg4(hc + o), hh = ldg4(hx + o);
    st4(dU + o, tf32_rn(v.x * (1.0f - rr.x) * (1.0f - cc.x * cc.x)), tf32_rn(v.y * (1.0f - rr.y) * (1.0f - cc.y * cc.y)),
        tf32_rn(v.z * (1.0f - rr.z) * (1.0f - cc.z * cc.z)), tf32_rn(v.w * (1.0f - rr.w) * (1.0f - cc.w * cc.w)));
    st4(dG + row * 2 * D + D + q4, tf32_rn(v.x * (hh.x - cc.x) * rr.x * (1.0f - rr.x)), tf32_rn(v.y * (hh.y - cc.y) * rr.y * (1.0f - rr.y)),
        tf32_rn(v.z * (hh.z - cc.z) * rr.z * (1.0f - rr.z)), tf32_rn(v.w * (hh.w - cc.w) * rr.w * (1.0f - rr.w)));
    st4(dHr + o, v.x * rr.x, v.y * rr.y, v.z * rr.z, v.w * rr.w);
  }
  if (proj) {
    __syncthreads();
    for (int i = threadIdx.x; i < Cout * D; i += blockDim.x) atomicAdd(dwp + i, sh_w[i]);
    if (threadIdx.x < Cout) {
      float s = 0.f;
      for (int i = 0; i < 32; ++i) s += sh_do[i * Cout + threadIdx.x];
      atomicAdd(dbp + threadIdx.x, s);
    }
  }
}

// ---- host side ----------------------------------------------------------------------------------
// St: transposed TF32 supports [KS][N][ldS].  dV: [R][nhalf*HS] (TF32-rounded).  wall: folded weights (hi part)
// [KS+2][HS][nhalf*HS].  qsave: [KS*nhalf][R][HS].  dib: [R][16].
template <int HS, class Epi>
int launch_agcn_bwd(int N, int B, int KS, int ldS, int nhalf, const float* St, const float* dV, const float* wall, float* qsave,
                    float* dib, const Epi& epi, cudaStream_t st) {
  using C = CfgB<HS>;
  const int64_t R = (int64_t)N * B;
  const int O = nhalf * HS;
  CUtensorMap tST, tVB, tVA, tW, tWib;
  {
    uint64_t dims[4] = {(uint64_t)N, (uint64_t)N, (uint64_t)KS, 1};
    uint64_t str[3] = {(uint64_t)ldS * 4, (uint64_t)N * ldS * 4, (uint64_t)KS * N * ldS * 4};
    uint32_t box[4] = {32, BM, 1, 1};
    MCRN_TRY(encode_tensor_map(&tST, St, dims, str, box, false));
  }
  {
    uint64_t dims[4] = {(uint64_t)O, (uint64_t)B, (uint64_t)N, 1};
    uint64_t str[3] = {(uint64_t)O * 4, (uint64_t)B * O * 4, (uint64_t)R * O * 4};
    uint32_t boxb[4] = {32, 1, BK, 1};
    MCRN_TRY(encode_tensor_map(&tVB, dV, dims, str, boxb, true));
    uint32_t boxa[4] = {32, 1, BM, 1};
    MCRN_TRY(encode_tensor_map(&tVA, dV, dims, str, boxa, false));
  }
  {
    uint64_t dims[4] = {(uint64_t)O, (uint64_t)HS, (uint64_t)(KS + 2), 1};
    uint64_t str[3] = {(uint64_t)O * 4, (uint64_t)HS * O * 4, (uint64_t)(KS + 2) * HS * O * 4};
    uint32_t box[4] = {32, (uint32_t)HS, 1, 1};
    MCRN_TRY(encode_tensor_map(&tW, wall, dims, str, box, false));
    uint32_t boxi[4] = {32, IBW, 1, 1};
    MCRN_TRY(encode_tensor_map(&tWib, wall, dims, str, boxi, false));
  }
  BParams p;
  p.N = N; p.B = B; p.KS = KS; p.nhalf = nhalf;
  p.qsave = qsave; p.blk_stride = R * HS; p.dib = dib;
  auto kern = agcn_bwd_kernel<HS, Epi>;
  static bool attr_set = false;
  if (!attr_set) {
    MCRN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    attr_set = true;
  }
  dim3 grid(ceil_div(N, BM), B, 1);
  const int pi = fused::prof_begin(fused::prof_class(1, HS, nhalf == 2 ? 1 : 0), st);
  MCRN_LAUNCH(kern, grid, BTHREADS, C::SMEM, st, tST, tVB, tVA, tW, tWib, p, epi);
  fused::prof_end(pi, st);
  return MCRN_OK;
}

}  // namespace fusedb
}  // namespace mcrn
