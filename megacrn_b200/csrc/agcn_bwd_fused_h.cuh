// Fused AGCN backward, fp16-operand version (default where the hidden width is 64 / 128).  Same algebra and structure
// as agcn_bwd_fused.cuh --
//     Q_ks[128 x HS] = S_k^T[tile rows, :] * dV[:, b, half]   (MMA1, fp32 accumulator in TMEM)
//     Q_ks -> fp16 pairs packed in place in TMEM; the fp32 value is stored for dW_k = sum X^T Q_k
//     acc[128 x HS] += Q_ks * W_k[:, half]^T (A from TMEM)  + dV_tile * W_0^T ;  accIB[128 x 16] = dV_tile * W_NB[0..16)^T
// -- with every tensor-core operand in HALF precision (kind::f16, fp32 accumulate): half the shared-memory fill bytes and
// twice the MMA rate of the TF32 version.  Gradients do not fit fp16's range as they are, so every gradient OPERAND is
// stored multiplied by one power-of-two loss scale `gs[0]` chosen per backward from max|d_output|, max|d_query|
// (k_grad_amax / k_grad_scale): the copies dV16 = fp16(dV * s) feed the MMAs, accumulators are multiplied by gs[1] = 1/s
// in the epilogues, and every fp32 tensor that leaves the kernels (dG, Q, dh_part, dH, dIB) is unscaled.  fp16 keeps
// the 11-bit significand of the TF32 path; the power-of-two scale changes no mantissa.
// All operands K-major, 128-byte swizzle:
//     S16T [KS][N][ld16]  transposed supports      A of MMA1   box [64 k][128 m]
//     dV16 [R][O]         rows (node, b)           B of MMA1, MN-major: HS/64 boxes [64 n][1][64 k] at column half * HS  (see agcn_fused_h.cuh)
//     dV16 [R][O]         rows                     A of MMA2 (identity segment)      box [64 k][1][128 m]
//     W16n [KS+2][HS][O]  folded weights, fp16     B of MMA2   box [64 k][HS n] and [64 k][16 n] (input block rows)
#pragma once

#include "agcn_bwd_fused.cuh"
#include "agcn_fused_h.cuh"

namespace mcrn {
namespace fusedbh {

using namespace tc;
using fused::mbar_arrive;
using fused::mbar_wait_b;
using fused::pow2_cols;
using fused::tmem_wait_st;
using fusedb::B_ITEM_P;
using fusedb::B_ITEM_SS;
using fusedb::B_ITEM_TS;
using fusedb::IBW;
using fusedb::tmem_ld_32x32b_x16;
using fusedh::BKH;
using fusedh::make_idesc_f16;
using fusedh::pack_h2;
using fusedh::round_h;
using fusedh::tcgen05_mma_f16;
using fusedh::tcgen05_mma_f16_ts;
using fusedh::tmem_st_32x32b_x16;

constexpr int BH_EPI_WARPS = 8;
constexpr int NPROD_B = 3;                      // TMA producer warps (see fusedh::NPROD): warp 0 and the last NPROD_B-1 warps
constexpr int BHTHREADS = 64 + 32 * BH_EPI_WARPS + 32 * (NPROD_B - 1);

struct BHParams {
  int N, B, KS, nhalf;
  float* qsave;          // [KS * nhalf][R][HS] fp32, unscaled (TF32 weight-gradient GEMMs); null when q16T is used
  int64_t blk_stride;    // R * HS
  __half* q16T;          // [KS * nhalf][R][HS] scaled fp16, ROW-major (operand of the fp16 weight-gradient kernel, read MN-major); or null
  float* dib;            // [R][IBW], unscaled (written when dxpin == null)
  const float* gs;       // device: {scale, 1 / scale}
  // gate-AGCN launch: the input-block gradient of the update AGCN (dib_in, written by the previous launch) is added and the
  // sum is scattered straight into dXPin [NB][R][Cin] (TF32-rounded), replacing a separate repack kernel
  const float* dib_in;
  float* dxpin;
  int nb, cin;
  int pdl_late;          // see fusedh::HParams
  unsigned long long* span;
};

template <int HS>
struct CfgBH {
  static_assert(HS == 64 || HS == 128, "hidden width: 64 or 128");
  static constexpr uint32_t A_SLOT = BM * 128;                    // [128 rows][64 halves]
  static constexpr uint32_t B_SLOT = (uint32_t)HS * 128;          // [HS rows][64 halves]
  static constexpr uint32_t IB_SLOT = IBW * 128;                  // [16 rows][64 halves]
  static constexpr uint32_t STAGE = A_SLOT + B_SLOT + IB_SLOT;
  static constexpr int NST = HS >= 128 ? 5 : 6;
  static constexpr uint32_t SCRATCH = 8 * 32 * 36 * 4;     // (rounding-phase staging of the fp32 Q path)
  static constexpr size_t SMEM = (size_t)NST * STAGE + SCRATCH + 1024;
  static constexpr uint32_t TM_ACC = 0, TM_IB = HS, TM_Q0 = HS + 32, TM_Q1 = 2 * HS + 32;
  static constexpr uint32_t TMEM_COLS = pow2_cols(3 * HS + 32);
  static constexpr int KB2 = HS / BKH;
};

// ---- epilogue functors: load4 / fin4 on 4 consecutive columns of one (node, b) row; acc is already unscaled ----------

// Update-AGCN tail (the cell's gate backward): dZH = acc; dG[:, :H] = dZH*h*z(1-z); dh_part = dHr + dZH*z.
struct EpiBUH {
  static constexpr int NP = 3;
  int H;
  const float *z, *h, *dHr;
  float *dG, *dh_part;     // dG: fp32 [R][2H] (columns [0, H) written here), or null when only the fp16 copy is consumed
  __half* g16;             // [R][2H] scaled fp16 copy (columns [0, H)): operand of the gate-AGCN launch
  __device__ __forceinline__ void load4(int row, int, int, int n0, float4 (&p)[NP]) const {
    const int64_t f = (int64_t)row * H + n0;
    p[0] = ldg4(z + f); p[1] = ldg4(h + f); p[2] = ldg4(dHr + f);
  }
  __device__ __forceinline__ void fin4(int row, int, int, int n0, const float4 (&p)[NP], const float (&acc)[4], float s, float inv_s) const {
    float st[1][4];
    const float4 zz = p[0], hh = p[1], dr = p[2];
    const float d0 = acc[0], d1 = acc[1], d2 = acc[2], d3 = acc[3];
    st[0][0] = round_h(d0 * hh.x * zz.x * (1.0f - zz.x) * s); st[0][1] = round_h(d1 * hh.y * zz.y * (1.0f - zz.y) * s);
    st[0][2] = round_h(d2 * hh.z * zz.z * (1.0f - zz.z) * s); st[0][3] = round_h(d3 * hh.w * zz.w * (1.0f - zz.w) * s);
    if (dG) st4(dG + (int64_t)row * 2 * H + n0, st[0][0] * inv_s, st[0][1] * inv_s, st[0][2] * inv_s, st[0][3] * inv_s);
    *reinterpret_cast<uint2*>(g16 + (int64_t)row * 2 * H + n0) = make_uint2(pack_h2(st[0][0], st[0][1]), pack_h2(st[0][2], st[0][3]));
    st4(dh_part + (int64_t)row * H + n0, dr.x + d0 * zz.x, dr.y + d1 * zz.y, dr.z + d2 * zz.z, dr.w + d3 * zz.w);
  }
};
// Gate-AGCN tail: dH_prev = acc + dh_part
struct EpiBGH {
  static constexpr int NP = 1;
  int H;
  const float* dh_part;
  float* dH_out;
  __device__ __forceinline__ void load4(int row, int, int, int n0, float4 (&p)[NP]) const { p[0] = ldg4(dh_part + (int64_t)row * H + n0); }
  __device__ __forceinline__ void fin4(int row, int, int, int n0, const float4 (&p)[NP], const float (&acc)[4], float, float) const {
    st4(dH_out + (int64_t)row * H + n0, acc[0] + p[0].x, acc[1] + p[0].y, acc[2] + p[0].z, acc[3] + p[0].w);
  }
};
// Gate-AGCN tail + the step glue of the NEXT cell to be processed (time step t-1), when that step's decoder input was
// teacher-forced (no gradient arrives through go) or in the encoder:
//   dH' = acc + dh_part + d_out_{t-1} . wp ;  dU' = dH'(1-r')(1-hc'^2) ;  dG'[:, H:] = dH'(h'-hc')r'(1-r') ;  dHr' = dH' r'
// with r', hc', h' the saved activations of step t-1.  Replaces k_bwd_glue_h for that step (dwp / dbp: k_proj_wgrad).
struct EpiBGHG {
  static constexpr int NP = 4;
  int H;
  const float *dh_part, *r, *hc, *hx;      // r, hc, hx: step t-1
  const float* dOut;                       // [B][T][N][Cout] upstream gradient or null
  const float* wp;                         // [Cout][H]
  int B, T, N, Cout, tprev;
  float *dU, *dG, *dHr;                    // step t-1: dU [R][H], dG [R][2H] (columns [H, 2H)), dHr [R][H]
  __half *u16, *g16;                       // scaled fp16 operand copies of step t-1 (row-major)
  __device__ __forceinline__ void load4(int row, int, int, int n0, float4 (&p)[NP]) const {
    const int64_t f = (int64_t)row * H + n0;
    p[0] = ldg4(dh_part + f); p[1] = ldg4(r + f); p[2] = ldg4(hc + f); p[3] = ldg4(hx + f);
  }
  __device__ __forceinline__ void fin4(int row, int node, int b, int n0, const float4 (&p)[NP], const float (&acc)[4], float s,
                                       float inv_s) const {
    float st[2][4];
    const float4 rr = p[1], cc = p[2], hh = p[3];
    float4 v = make_float4(acc[0] + p[0].x, acc[1] + p[0].y, acc[2] + p[0].z, acc[3] + p[0].w);
    if (dOut != nullptr) {
      for (int co = 0; co < Cout; ++co) {
        const float d = __ldg(dOut + (((int64_t)b * T + tprev) * N + node) * Cout + co);
        const float4 w4 = ldg4(wp + (int64_t)co * H + n0);
        v.x = fmaf(d, w4.x, v.x); v.y = fmaf(d, w4.y, v.y); v.z = fmaf(d, w4.z, v.z); v.w = fmaf(d, w4.w, v.w);
      }
    }
    st[0][0] = round_h(v.x * (1.0f - rr.x) * (1.0f - cc.x * cc.x) * s); st[0][1] = round_h(v.y * (1.0f - rr.y) * (1.0f - cc.y * cc.y) * s);
    st[0][2] = round_h(v.z * (1.0f - rr.z) * (1.0f - cc.z * cc.z) * s); st[0][3] = round_h(v.w * (1.0f - rr.w) * (1.0f - cc.w * cc.w) * s);
    st[1][0] = round_h(v.x * (hh.x - cc.x) * rr.x * (1.0f - rr.x) * s); st[1][1] = round_h(v.y * (hh.y - cc.y) * rr.y * (1.0f - rr.y) * s);
    st[1][2] = round_h(v.z * (hh.z - cc.z) * rr.z * (1.0f - rr.z) * s); st[1][3] = round_h(v.w * (hh.w - cc.w) * rr.w * (1.0f - rr.w) * s);
    const int64_t o = (int64_t)row * H + n0, og = (int64_t)row * 2 * H + H + n0;
    st4(dU + o, st[0][0] * inv_s, st[0][1] * inv_s, st[0][2] * inv_s, st[0][3] * inv_s);
    st4(dG + og, st[1][0] * inv_s, st[1][1] * inv_s, st[1][2] * inv_s, st[1][3] * inv_s);
    st4(dHr + o, v.x * rr.x, v.y * rr.y, v.z * rr.z, v.w * rr.w);
    *reinterpret_cast<uint2*>(u16 + o) = make_uint2(pack_h2(st[0][0], st[0][1]), pack_h2(st[0][2], st[0][3]));
    *reinterpret_cast<uint2*>(g16 + og) = make_uint2(pack_h2(st[1][0], st[1][1]), pack_h2(st[1][2], st[1][3]));
  }
};

template <int HS, class Epi>
__global__ void __launch_bounds__(BHTHREADS, 1)
agcn_bwd_h_kernel(const __grid_constant__ CUtensorMap tmST, const __grid_constant__ CUtensorMap tmVB,
                  const __grid_constant__ CUtensorMap tmVA, const __grid_constant__ CUtensorMap tmW,
                  const __grid_constant__ CUtensorMap tmWib, BHParams p, Epi epi) {
  using C = CfgBH<HS>;
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
  const int kb1 = (p.N + BKH - 1) / BKH;
  const int nks = p.KS * p.nhalf;
  const int NBLK = p.KS + 1;
  if (p.span != nullptr && threadIdx.x == 0) atomicMin(p.span, fused::globaltimer_ns());

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
  pdl_wait();                       // programmatic dependent launch: the prologue above overlapped the previous kernel's tail
  if (!p.pdl_late) pdl_launch_dependents();

  const int pw = warp == 0 ? 0 : (warp >= 2 + BH_EPI_WARPS ? warp - (2 + BH_EPI_WARPS) + 1 : -1);   // producer index or -1
  if (pw >= 0) {
    if (lane == 0) {                                     // ===== TMA producers: item `it` belongs to producer it % NPROD_B =====
      int it = 0;
      fusedb::for_each_item_b<C::KB2>(nks, p.nhalf, kb1, [&](int type, int ks, int j) {
        if (it % NPROD_B != pw) { ++it; return; }
        const int s = it % NST;
        if (it >= NST) mbar_wait_b(smem_u32(&empty_bar[s]), (((uint32_t)(it / NST)) & 1u) ^ 1u);
        const uint32_t fb = smem_u32(&full_bar[s]);
        const uint32_t a_dst = smem_base + (uint32_t)s * C::STAGE, b_dst = a_dst + C::A_SLOT, ib_dst = b_dst + C::B_SLOT;
        if (type == B_ITEM_P) {
          const int k = ks / p.nhalf, half = ks - k * p.nhalf;
          mbar_expect_tx(fb, C::A_SLOT + C::B_SLOT);
          tma_load_4d(a_dst, &tmST, fb, j * BKH, m0, k, 0);                        // S_k^T[m0.., 64 j..]
#pragma unroll
          for (int q = 0; q < HS / 64; ++q)                                        // dV rows (64 j.., b), columns half*HS + 64 q..  (MN-major B)
            tma_load_4d(b_dst + q * 8192, &tmVB, fb, half * HS + q * 64, b, j * BKH, 0);
        } else if (type == B_ITEM_SS) {
          const int half = ks;
          mbar_expect_tx(fb, C::A_SLOT + C::B_SLOT + C::IB_SLOT);
          tma_load_4d(a_dst, &tmVA, fb, half * HS + j * BKH, b, m0, 0);            // dV rows (m0.., b), columns (o) 64 j..
          tma_load_4d(b_dst, &tmW, fb, half * HS + j * BKH, 0, 0, 0);              // W_0[0..HS][o-block]
          tma_load_4d(ib_dst, &tmWib, fb, half * HS + j * BKH, 0, NBLK, 0);        // W_NB[0..16][o-block]
        } else {
          const int k = ks / p.nhalf, half = ks - k * p.nhalf;
          mbar_expect_tx(fb, C::B_SLOT);
          tma_load_4d(b_dst, &tmW, fb, half * HS + j * BKH, 0, 1 + k, 0);          // W_{1+k}[0..HS][o-block]
        }
        ++it;
      });
    }
  } else if (warp == 1) {
    if (lane == 0) {                                     // ===== MMA issuer =====
      constexpr uint32_t idesc_n = make_idesc_f16<HS>();
      constexpr uint32_t idesc_p = make_idesc_f16<HS>() | (1u << 16);      // MMA1: B is MN-major (row-major dV rows)
      constexpr uint32_t idesc_ib = make_idesc_f16<IBW>();
      int it = 0;
      bool acc_on = false, ib_on = false;
      fusedb::for_each_item_b<C::KB2>(nks, p.nhalf, kb1, [&](int type, int ks, int j) {
        const int s = it % NST;
        mbar_wait_b(smem_u32(&full_bar[s]), ((uint32_t)(it / NST)) & 1u);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_base + (uint32_t)s * C::STAGE, b_addr = a_addr + C::A_SLOT, ib_addr = b_addr + C::B_SLOT;
        const uint32_t qbuf = tmem_base + ((ks & 1) ? C::TM_Q1 : C::TM_Q0);
        if (type == B_ITEM_P) {
          const int nkk = min(BKH / 16, (p.N - j * BKH + 15) / 16);     // the last k-block stops at the node count
#pragma unroll
          for (int kk = 0; kk < BKH / 16; ++kk) {
            if (kk < nkk) {
              const uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024, 2);
              const uint64_t bd = make_smem_desc(b_addr + kk * 2048, 8192, 1024, 2);
              tcgen05_mma_f16(qbuf, ad, bd, idesc_p, (j > 0 || kk > 0) ? 1u : 0u);
            }
          }
          tcgen05_commit(smem_u32(&empty_bar[s]));
          if (j == kb1 - 1) tcgen05_commit(smem_u32(&q_full_bar[ks & 1]));
        } else if (type == B_ITEM_SS) {
#pragma unroll
          for (int kk = 0; kk < BKH / 16; ++kk) {
            const uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024, 2);
            const uint64_t bd = make_smem_desc(b_addr + kk * 32, 16, 1024, 2);
            const uint64_t id = make_smem_desc(ib_addr + kk * 32, 16, 1024, 2);
            tcgen05_mma_f16(tmem_base + C::TM_ACC, ad, bd, idesc_n, (acc_on || kk > 0) ? 1u : 0u);
            tcgen05_mma_f16(tmem_base + C::TM_IB, ad, id, idesc_ib, (ib_on || kk > 0) ? 1u : 0u);
          }
          acc_on = true; ib_on = true;
          tcgen05_commit(smem_u32(&empty_bar[s]));
        } else {
          if (j == 0) {
            mbar_wait_b(smem_u32(&q_ready_bar[ks & 1]), ((uint32_t)(ks >> 1)) & 1u);
            tcgen05_fence_after();
          }
#pragma unroll
          for (int kk = 0; kk < BKH / 16; ++kk) {
            const uint64_t bd = make_smem_desc(b_addr + kk * 32, 16, 1024, 2);
            tcgen05_mma_f16_ts(tmem_base + C::TM_ACC, qbuf + (uint32_t)(j * (BKH / 2) + kk * 8), bd, idesc_n, (acc_on || kk > 0) ? 1u : 0u);
          }
          acc_on = true;
          tcgen05_commit(smem_u32(&empty_bar[s]));
        }
        ++it;
      });
      tcgen05_commit(smem_u32(&acc_full_bar));
    }
  } else if (warp < 2 + BH_EPI_WARPS) {                  // ===== rounding + epilogue warps =====
    const int quarter = warp & 3;
    const int ew = warp - 2, half_id = ew >> 2;
    const int cq = (lane & 7) * 4, r0 = lane >> 3;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int node0 = m0 + quarter * 32;
    const float gs = __ldg(p.gs), inv_gs = __ldg(p.gs + 1);
    {
      // Q_ks: fp32 accumulator -> fp16 pairs packed in place; the unscaled fp32 value is stored for the weight gradient.
      // The two warps of a lane quarter take alternate 32-column chunks; chunk c packs into columns [16c, 16c+16), which
      // chunk c' < c of the OTHER warp may not have read yet: all reads of the buffer happen before any packed store.
      float* scr = reinterpret_cast<float*>(smem_al + (size_t)NST * C::STAGE) + ew * (32 * 36);
      for (int ks = 0; ks < nks; ++ks) {
        mbar_wait_b(smem_u32(&q_full_bar[ks & 1]), ((uint32_t)(ks >> 1)) & 1u);
        tcgen05_fence_after();
        const uint32_t qbuf = tmem_base + ((ks & 1) ? C::TM_Q1 : C::TM_Q0) + lane_off;
        constexpr int NCH = (HS / 32 + 1) / 2;             // chunks per warp
        uint32_t u[NCH][16];
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = half_id + 2 * ci;
          if (c < HS / 32) {
            float v[32];
            tmem_ld_32x32b_x32(qbuf + (uint32_t)(c * 32), v);
#pragma unroll
            for (int i = 0; i < 16; ++i) u[ci][i] = pack_h2(v[2 * i], v[2 * i + 1]);
            if (p.q16T != nullptr) {                     // row-major fp16 copy straight from the accumulator layout (lane = node row):
              // 32 consecutive columns = 64 contiguous bytes per lane
              const int node = node0 + lane;
              if (node < p.N) {
                uint4* dst = reinterpret_cast<uint4*>(p.q16T + ((int64_t)ks * p.N * p.B + (int64_t)node * p.B + b) * HS + c * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = make_uint4(u[ci][4 * i], u[ci][4 * i + 1], u[ci][4 * i + 2], u[ci][4 * i + 3]);
              }
            } else if (node0 < p.N) {
              __syncwarp();
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4*>(&scr[lane * 36 + i]) = make_float4(round_h(v[i]) * inv_gs, round_h(v[i + 1]) * inv_gs,
                                                                              round_h(v[i + 2]) * inv_gs, round_h(v[i + 3]) * inv_gs);
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
        }
        // both warps of this quarter have finished reading the fp32 buffer (named barrier per quarter: 64 threads)
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
#pragma unroll
        for (int ci = 0; ci < NCH; ++ci) {
          const int c = half_id + 2 * ci;
          if (c < HS / 32) tmem_st_32x32b_x16(qbuf + (uint32_t)(c * 16), u[ci]);
        }
        tmem_wait_st();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&q_ready_bar[ks & 1]));
      }
    }
    mbar_wait_b(smem_u32(&acc_full_bar), 0);
    tcgen05_fence_after();
    if (p.pdl_late) pdl_launch_dependents();
    if (node0 < p.N) {
      float* scr = reinterpret_cast<float*>(smem_al) + ew * (32 * 36);      // the ring is idle now
#pragma unroll 1
      for (int c = half_id; c < HS / 32; c += 2) {
        float v[32];
        tmem_ld_32x32b_x32(tmem_base + C::TM_ACC + lane_off + (uint32_t)(c * 32), v);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(&scr[lane * 36 + i]) = make_float4(v[i] * inv_gs, v[i + 1] * inv_gs, v[i + 2] * inv_gs, v[i + 3] * inv_gs);
        __syncwarp();
        const int col = c * 32 + cq;
        constexpr int RB = Epi::NP <= 2 ? 8 : 4;
#pragma unroll
        for (int b0 = 0; b0 < 8; b0 += RB) {
          float4 pre[RB][Epi::NP];
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const int node = node0 + r0 + 4 * (b0 + i);
            if (node < p.N) epi.load4(node * p.B + b, node, b, col, pre[i]);
          }
#pragma unroll
          for (int i = 0; i < RB; ++i) {
            const int rr = r0 + 4 * (b0 + i), node = node0 + rr;
            if (node < p.N) {
              const float4 t = *reinterpret_cast<const float4*>(&scr[rr * 36 + cq]);
              const float a4[4] = {t.x, t.y, t.z, t.w};
              epi.fin4(node * p.B + b, node, b, col, pre[i], a4, gs, inv_gs);
            }
          }
        }
        __syncwarp();                                    // the staging tile is reused by the next chunk
      }
      if (half_id == 1) {                                // input-block gradient: 16 columns, one row per thread
        float v[16];
        tmem_ld_32x32b_x16(tmem_base + C::TM_IB + lane_off, v);
        const int node = node0 + lane;
        if (node < p.N) {
          const int64_t row = (int64_t)node * p.B + b;
          if (p.dxpin == nullptr) {
            float* dst = p.dib + row * IBW;
#pragma unroll
            for (int i = 0; i < 16; i += 4) st4(dst + i, v[i] * inv_gs, v[i + 1] * inv_gs, v[i + 2] * inv_gs, v[i + 3] * inv_gs);
          } else {
            const float* src = p.dib_in + row * IBW;
            const int64_t R = (int64_t)p.N * p.B;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 u4 = ldg4(src + i);
              v[i] = v[i] * inv_gs + u4.x; v[i + 1] = v[i + 1] * inv_gs + u4.y; v[i + 2] = v[i + 2] * inv_gs + u4.z; v[i + 3] = v[i + 3] * inv_gs + u4.w;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (j < p.nb * p.cin) {
                const int k = j / p.cin, ci = j - k * p.cin;
                p.dxpin[((int64_t)k * R + row) * p.cin + ci] = tf32_rn(v[j]);
              }
            }
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
  if (p.span != nullptr && threadIdx.x == 0) atomicMax(p.span + 1, fused::globaltimer_ns());
}

// ---- loss scale ---------------------------------------------------------------------------------
// amax[0] = max |x| over all upstream gradient tensors (as the bit pattern of a non-negative float: integer max is float max)
struct AmaxSrc { const float* p[5]; int64_t end[5]; };     // end[i] = cumulative element count
__global__ void __launch_bounds__(256) k_grad_amax(AmaxSrc a, unsigned* __restrict__ amax) {
  __shared__ float sh[8];
  float m = 0.f;
  const int64_t total = a.end[4];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int k = 0;
    while (i >= a.end[k]) ++k;
    m = fmaxf(m, fabsf(a.p[k][i - (k ? a.end[k - 1] : 0)]));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 8) {                          // one atomic per block (the old one-per-warp version spent 20 us in contention)
    m = sh[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffu, m, o));
    if (threadIdx.x == 0 && m > 0.f && m < INFINITY) atomicMax(amax, __float_as_uint(m));
  }
}
// gs = {2^e, 2^-e} with amax * 2^e in [2^3, 2^4): x4096-8192 of head-room below fp16's 65504 for the gradients BPTT builds up
// from the upstream ones (amax is taken over the UPSTREAM gradients only), full fp16 precision down to 2^-18 amax and
// subnormals down to 2^-28 amax.  (Round 2 first used [2^7, 2^8): a real METR-LA run hit an fp16 overflow -> NaN after 21 k
// steps, profiles/r2_metrla_real_run.txt.)  Beyond the head-room the operand conversions saturate (fusedh::pack_h2 / round_h).
__global__ void k_grad_scale(const unsigned* __restrict__ amax, float* __restrict__ gs) {
  const float m = __uint_as_float(*amax);
  int e = 0;
  if (m > 0.f) {
    int ex;
    frexpf(m, &ex);                 // m = f * 2^ex, f in [0.5, 1)
    e = 4 - ex;
    e = e > 60 ? 60 : (e < -60 ? -60 : e);
  }
  gs[0] = exp2f((float)e);
  gs[1] = exp2f((float)-e);
}

// ---- operand conversions ------------------------------------------------------------------------
// S16T[k][m][n] = fp16(S[k][n][m])
__global__ void k_supports_to_half_T(const float* __restrict__ S, __half* __restrict__ St, int n, int ld, int ld16) {
  __shared__ float tile[32][33];
  const int k = blockIdx.z;
  const float* s = S + (int64_t)k * n * ld;
  __half* t = St + (int64_t)k * n * ld16;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = by + i, c = bx + threadIdx.x;
    tile[i][threadIdx.x] = (r < n && c < n) ? s[(int64_t)r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = bx + i, c = by + threadIdx.x;
    if (r < n && c < ld16) t[(int64_t)r * ld16 + c] = __float2half_rn(c < n ? tile[threadIdx.x][i] : 0.f);
  }
}
// folded weights fp32 [2 (TF32 hi, lo)][nseg][HS][O] -> fp16 [nseg][HS][O] (same layout; the backward uses one part)
__global__ void k_weights_to_half_n(const float* __restrict__ wall, __half* __restrict__ w16, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    w16[i] = __float2half_rn(wall[i] + wall[total + i]);
}

// ---- step glue, fp16 version ----------------------------------------------------------------------
// As fusedb::k_bwd_glue, plus the scaled fp16 operand copies (row-major) of dU and of the r-half of dG.  Block = one batch element b x 32 consecutive nodes (rows n*B + b), thread = 4 columns.
__global__ void __launch_bounds__(256) k_bwd_glue_h(const float* __restrict__ dOut, const float* __restrict__ dxin, int dxin_stride,
                                                    const float* __restrict__ h_t, const float* __restrict__ wp,
                                                    const float* __restrict__ dH, int dh_init, const float* __restrict__ r,
                                                    const float* __restrict__ hc, const float* __restrict__ hx,
                                                    float* __restrict__ dU, float* __restrict__ dG, float* __restrict__ dHr,
                                                    __half* __restrict__ u16, __half* __restrict__ g16, const float* __restrict__ gsp,
                                                    float* __restrict__ dwp, float* __restrict__ dbp, int B, int T, int N,
                                                    int D, int Cout, int t) {
  extern __shared__ float sh[];                    // [32][Cout] d_out rows, [Cout][D] dwp partials
  pdl_wait();
  pdl_launch_dependents();
  float* sh_do = sh;
  float* sh_w = sh + 32 * Cout;
  const int b = blockIdx.y, n0 = blockIdx.x * 32;
  const bool proj = (dOut != nullptr) || (dxin != nullptr);
  const float s = __ldg(gsp), inv_s = __ldg(gsp + 1);
  if (proj) {
    for (int i = threadIdx.x; i < 32 * Cout; i += blockDim.x) {
      const int n = n0 + i / Cout, co = i % Cout;
      float v = 0.f;
      if (n < N) {
        if (dOut) v = dOut[(((int64_t)b * T + t) * N + n) * Cout + co];
        if (dxin) v += dxin[((int64_t)n * B + b) * dxin_stride + co];
      }
      sh_do[i] = v;
    }
    for (int i = threadIdx.x; i < Cout * D; i += blockDim.x) sh_w[i] = 0.f;
    __syncthreads();
  }
  const int Q = D >> 2;
  for (int e = threadIdx.x; e < 32 * Q; e += blockDim.x) {
    const int i = e / Q, q4 = (e - i * Q) * 4;
    const int n = n0 + i;
    if (n >= N) continue;
    const int64_t row = (int64_t)n * B + b;
    const int64_t o = row * D + q4;
    float4 v = dh_init ? make_float4(0.f, 0.f, 0.f, 0.f) : ldg4(dH + o);
    if (proj) {
      // dwp == null: the projection weight gradient of this step is computed off the chain (k_proj_wgrad)
      const float4 hv = dwp ? ldg4(h_t + o) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int co = 0; co < Cout; ++co) {
        const float d = sh_do[i * Cout + co];
        const float4 w4 = ldg4(wp + (int64_t)co * D + q4);
        v.x = fmaf(d, w4.x, v.x); v.y = fmaf(d, w4.y, v.y); v.z = fmaf(d, w4.z, v.z); v.w = fmaf(d, w4.w, v.w);
        if (dwp) {
          float* sw = sh_w + co * D + q4;
          atomicAdd(sw, d * hv.x); atomicAdd(sw + 1, d * hv.y); atomicAdd(sw + 2, d * hv.z); atomicAdd(sw + 3, d * hv.w);
        }
      }
    }
    const float4 rr = ldg4(r + o), cc = ldg4(hc + o), hh = ldg4(hx + o);
    // scaled, fp16-rounded operand values; the fp32 tensors hold exactly value / scale
    const float u0 = round_h(v.x * (1.0f - rr.x) * (1.0f - cc.x * cc.x) * s), u1 = round_h(v.y * (1.0f - rr.y) * (1.0f - cc.y * cc.y) * s);
    const float u2 = round_h(v.z * (1.0f - rr.z) * (1.0f - cc.z * cc.z) * s), u3 = round_h(v.w * (1.0f - rr.w) * (1.0f - cc.w * cc.w) * s);
    const float g0 = round_h(v.x * (hh.x - cc.x) * rr.x * (1.0f - rr.x) * s), g1 = round_h(v.y * (hh.y - cc.y) * rr.y * (1.0f - rr.y) * s);
    const float g2 = round_h(v.z * (hh.z - cc.z) * rr.z * (1.0f - rr.z) * s), g3 = round_h(v.w * (hh.w - cc.w) * rr.w * (1.0f - rr.w) * s);
    if (dU) {                                        // fp32 copies: only the TF32 weight- / support-gradient paths read them
      st4(dU + o, u0 * inv_s, u1 * inv_s, u2 * inv_s, u3 * inv_s);
      st4(dG + row * 2 * D + D + q4, g0 * inv_s, g1 * inv_s, g2 * inv_s, g3 * inv_s);
    }
    st4(dHr + o, v.x * rr.x, v.y * rr.y, v.z * rr.z, v.w * rr.w);
    *reinterpret_cast<uint2*>(u16 + o) = make_uint2(pack_h2(u0, u1), pack_h2(u2, u3));
    *reinterpret_cast<uint2*>(g16 + row * 2 * D + D + q4) = make_uint2(pack_h2(g0, g1), pack_h2(g2, g3));
  }
  __syncthreads();
  if (proj && dwp) {
    for (int i = threadIdx.x; i < Cout * D; i += blockDim.x) atomicAdd(dwp + i, sh_w[i]);
    if (threadIdx.x < Cout) {
      float sm = 0.f;
      for (int i = 0; i < 32; ++i) sm += sh_do[i * Cout + threadIdx.x];
      atomicAdd(dbp + threadIdx.x, sm);
    }
  }
}

// Projection weight gradient of the steps whose glue ran inside the gate-AGCN epilogue (bit t of `mask`):
//   dwp[co][:] += sum_rows d_out_t[row][co] * h_t[row][:],  dbp[co] += sum_rows d_out_t[row][co]
// h_t = hx_base + (t + 1) * hx_step for t + 1 < T, else h_last.  Grid (row blocks of 32, T); thread = 4 columns.
__global__ void __launch_bounds__(256) k_proj_wgrad(const float* __restrict__ dOut, const float* __restrict__ hx_base, int64_t hx_step,
                                                    const float* __restrict__ h_last, unsigned mask, float* __restrict__ dwp,
                                                    float* __restrict__ dbp, int B, int T, int N, int D, int Cout) {
  const int t = blockIdx.y;
  if (!((mask >> t) & 1u)) return;
  extern __shared__ float sh[];                    // [32][Cout] d_out rows, [Cout][D] partials
  float* sh_do = sh;
  float* sh_w = sh + 32 * Cout;
  const int64_t R = (int64_t)N * B, r0 = (int64_t)blockIdx.x * 32;
  const float* h_t = (t + 1 < T) ? hx_base + (int64_t)(t + 1) * hx_step : h_last;
  for (int i = threadIdx.x; i < 32 * Cout; i += blockDim.x) {
    const int64_t row = r0 + i / Cout;
    const int co = i % Cout;
    float v = 0.f;
    if (row < R) {
      const int n = (int)(row / B), b = (int)(row % B);
      v = dOut[(((int64_t)b * T + t) * N + n) * Cout + co];
    }
    sh_do[i] = v;
  }
  for (int i = threadIdx.x; i < Cout * D; i += blockDim.x) sh_w[i] = 0.f;
  __syncthreads();
  const int Q = D >> 2;
  for (int e = threadIdx.x; e < 32 * Q; e += blockDim.x) {
    const int i = e / Q, q4 = (e - i * Q) * 4;
    const int64_t row = r0 + i;
    if (row >= R) continue;
    const float4 hv = ldg4(h_t + row * D + q4);
    for (int co = 0; co < Cout; ++co) {
      const float d = sh_do[i * Cout + co];
      float* sw = sh_w + co * D + q4;
      atomicAdd(sw, d * hv.x); atomicAdd(sw + 1, d * hv.y); atomicAdd(sw + 2, d * hv.z); atomicAdd(sw + 3, d * hv.w);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cout * D; i += blockDim.x) atomicAdd(dwp + i, sh_w[i]);
  if (threadIdx.x < Cout) {
    float sm = 0.f;
    for (int i = 0; i < 32; ++i) sm += sh_do[i * Cout + threadIdx.x];
    atomicAdd(dbp + threadIdx.x, sm);
  }
}

// ---- host side ----------------------------------------------------------------------------------
struct BHOperands {
  const __half* S16T;    // [KS][N][ld_half(N)]
  const __half* V16;     // [R][O]
  const __half* W16n;    // [KS+2][HS][O]
  const float* gs;       // device {scale, 1/scale}
};

template <int HS, class Epi>
int launch_agcn_bwd_h(int N, int B, int KS, int nhalf, const BHOperands& op, float* qsave, float* dib, const Epi& epi,
                      cudaStream_t st, const float* dib_in = nullptr, float* dxpin = nullptr, int cin = 0, __half* q16T = nullptr) {
  using C = CfgBH<HS>;
  const int64_t R = (int64_t)N * B;
  const int O = nhalf * HS, ldn = fusedh::ld_half(N);
  CUtensorMap tST, tVB, tVA, tW, tWib;
  {
    uint64_t dims[4] = {(uint64_t)N, (uint64_t)N, (uint64_t)KS, 1};
    uint64_t str[3] = {(uint64_t)ldn * 2, (uint64_t)N * ldn * 2, (uint64_t)KS * N * ldn * 2};
    uint32_t box[4] = {BKH, BM, 1, 1};
    MCRN_TRY(fusedh::encode_tensor_map_h(&tST, op.S16T, dims, str, box));
  }
  {
    uint64_t dims[4] = {(uint64_t)O, (uint64_t)B, (uint64_t)N, 1};
    uint64_t str[3] = {(uint64_t)O * 2, (uint64_t)B * O * 2, (uint64_t)R * O * 2};
    uint32_t box[4] = {BKH, 1, BM, 1};
    MCRN_TRY(fusedh::encode_tensor_map_h(&tVA, op.V16, dims, str, box));
    uint32_t boxb[4] = {64, 1, BKH, 1};                  // MN-major B of MMA1: 64 columns x 64 node rows of batch element b
    MCRN_TRY(fusedh::encode_tensor_map_h(&tVB, op.V16, dims, str, boxb));
  }
  {
    uint64_t dims[4] = {(uint64_t)O, (uint64_t)HS, (uint64_t)(KS + 2), 1};
    uint64_t str[3] = {(uint64_t)O * 2, (uint64_t)HS * O * 2, (uint64_t)(KS + 2) * HS * O * 2};
    uint32_t box[4] = {BKH, (uint32_t)HS, 1, 1};
    MCRN_TRY(fusedh::encode_tensor_map_h(&tW, op.W16n, dims, str, box));
    uint32_t boxi[4] = {BKH, IBW, 1, 1};
    MCRN_TRY(fusedh::encode_tensor_map_h(&tWib, op.W16n, dims, str, boxi));
  }
  BHParams p;
  p.N = N; p.B = B; p.KS = KS; p.nhalf = nhalf;
  p.qsave = qsave; p.blk_stride = R * HS; p.dib = dib; p.gs = op.gs;
  p.q16T = q16T;
  p.dib_in = dib_in; p.dxpin = dxpin; p.nb = KS + 1; p.cin = cin;
  p.pdl_late = (g_pdl_chain >> 3) & 1;
  p.span = fused::next_span();
  auto kern = agcn_bwd_h_kernel<HS, Epi>;
  static bool attr_set = false;
  if (!attr_set) {
    MCRN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    attr_set = true;
  }
  dim3 grid(ceil_div(N, BM), B, 1);
  const int pi = fused::prof_begin(fused::prof_class(1, HS, nhalf == 2 ? 1 : 0), st);
  MCRN_TRY(launch_chain(2, kern, grid, dim3(BHTHREADS), C::SMEM, st, "agcn_bwd_h_kernel", tST, tVB, tVA, tW, tWib, p, epi));
  fused::prof_end(pi, st);
  return MCRN_OK;
}

}  // namespace fusedbh
}  // namespace mcrn
