// Fused AGCN forward, fp16-operand version (the default forward path for hidden widths 64 / 128).
//
// Same structure as agcn_fused.cuh -- per CTA (128-node tile, batch element b):
//     MMA1   P_k[128 x HS] = S_k[tile rows, :] * X[:, b, :]                 (tcgen05.mma.kind::f16, fp32 accumulate, TMEM)
//     round  P_k -> fp16 pairs, packed IN PLACE in TMEM (tcgen05.ld / cvt.rn.f16x2 / tcgen05.st)
//     MMA2   acc[128 x O] += P_k * W_k   (A = packed P_k straight from TMEM)  + X_tile * W_0 + IB_tile * W_NB  (A in smem)
//     epilogue: sigmoid / tanh / GRU algebra (model/MegaCRN.py:43-47), writes the fp32 tensors the backward needs and the
//               row-major fp16 operand copy the next fused launch reads
// -- but every tensor-core operand is stored in HALF precision.  fp16 has the same 11-bit significand as TF32, so the
// numerics match the TF32 path (all forward operands are O(1): softmax supports, states in (-1,1), inputs, weights), while
// each operand byte carries twice the work: the kernel is bound by operand bytes through shared memory (every byte is
// written by TMA and read by the MMA at 128 B/clk/SM; profiles/r1_summary.md), and kind::f16 also runs at twice the
// kind::tf32 MMA rate.  The weights keep the hi + lo split
// (hi = fp16(W), lo = fp16(W - hi)).  All operands are K-major with the 128-byte swizzle:
//     S16  [KS][N][ld16]        supports                       A of MMA1   box [64 k][128 m]
//     X16  [R][HS]              state rows (node, b)           B of MMA1, MN-major: HS/64 boxes [64 n][1][64 k]  (no transposed copy:
//                                                              descriptor LBO 8192 / SBO 1024 / SWIZZLE_128B, K step 2048 B; tools/probe_mn16.py)
//     X16  [R][HS], IB16 [R][HS] state / input block rows      A of MMA2 (identity + input segments)  box [64 k][1][128 m]
//     W16  [parts][KS+2][O][HS] folded weights, transposed     B of MMA2   box [64 k][O n]
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = P rounding + epilogue,
// warps 6..17 = epilogue only (each TMEM lane quarter is served by four warps that split the column chunks).
#pragma once

#include <cuda_fp16.h>

#include "agcn_fused.cuh"

namespace mcrn {
namespace fusedh {

using namespace tc;
using fused::mbar_arrive;
using fused::mbar_wait_b;
using fused::pow2_cols;
using fused::tmem_wait_st;
using fused::ITEM_P;
using fused::ITEM_SS;
using fused::ITEM_TS;

constexpr int EPI_WARPS = 16;                   // epilogue warps: 4 per TMEM lane quarter (latency-bound tail: more warps in flight)
// TMA producer warps: one elected thread issues one cp.async.bulk.tensor per ~190 cycles (measured, tools/probe_tma2.cu:
// issue-bound, independent of the box size), so a single producer caps the ring at ~45 B/clk; the ring items are dealt
// round-robin to NPROD producer warps (warp 0 and the last NPROD-1 warps of the CTA).
constexpr int NPROD = 4;
constexpr int FTHREADS = 64 + 32 * EPI_WARPS + 32 * (NPROD - 1);
constexpr int BKH = 64;                         // halves per k-block = one 128-byte swizzle row

struct HParams {
  int N, B, KS, nparts;
  int ib_blocks;        // k-blocks (64 halves) of the input-block segment: HS/64, or 1 for the compact [R][64] input block
  float* xp_save;       // training: fp32 XP buffer [KS+2][R][HS]; the (fp16-rounded) P_k goes to block 1+k.  null = eval
  int64_t blk_stride;   // R * HS
  long long* dbg;       // debug timeline (see agcn_fused.cuh)
  unsigned long long* span;   // debug: {min CTA start, max CTA end} of this launch (ns), or null
  int pdl_late;         // programmatic dependent launch: 0 = let the dependents start right after the prologue, 1 = at the epilogue
};

__device__ __forceinline__ void tcgen05_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// Instruction descriptor, kind::f16: D = F32 (bits [4,6) = 1), A = B = F16 (format 0), both K-major, N>>3 [17,23), M>>4 [24,29).
template <int N_>
__device__ __forceinline__ constexpr uint32_t make_idesc_f16() {
  return (1u << 4) | ((uint32_t)(N_ >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// fp32 -> fp16, round to nearest, SATURATING to +-65504 (one F2FP.SATFINITE, same cost as the plain conversion): a scaled
// gradient operand that outgrows fp16's range (BPTT can amplify the recurrent gradient far beyond the upstream maximum the
// loss scale is chosen from) is clipped instead of becoming inf -- no inf, hence no inf - inf = NaN, can enter an MMA.
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t r;                                  // low 16 bits = lo
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float round_h(float v) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(v));
  return __half2float(__ushort_as_half(h));
}
// Gate activations: one ex2 and one rcp (MUFU) each, flush-to-zero forms -- no range-check code around them (the epilogue
// issues 2 x 32 K of these per tile).  ex2.approx: max rel. error 2^-22; |x| large: e -> 0 or +inf, rcp(inf) = 0: exact limits.
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_ftz(1.0f + ex2_ftz(-1.4426950408889634f * x)); }
__device__ __forceinline__ float tanh_fast(float x) {
  const float t = ex2_ftz(-2.8853900817779268f * fabsf(x));
  return copysignf((1.0f - t) * rcp_ftz(1.0f + t), x);
}

// ring items of one CTA, in issue order (as fused::for_each_item, but the input-block segment has `ibk` k-blocks)
template <int KB2, class F>
__device__ __forceinline__ void for_each_item_h(int KS, int kb1, int nparts, int ibk, F&& f) {
  for (int j = 0; j < kb1; ++j) f(ITEM_P, 0, j, 0);
  if (KS > 1)
    for (int j = 0; j < kb1; ++j) f(ITEM_P, 1, j, 0);
  for (int part = 0; part < nparts; ++part)
    for (int j = 0; j < KB2; ++j) f(ITEM_SS, 0, j, part);
  for (int part = 0; part < nparts; ++part)
    for (int j = 0; j < ibk; ++j) f(ITEM_SS, KS + 1, j, part);
  for (int k = 0; k < KS; ++k) {
    for (int part = 0; part < nparts; ++part)
      for (int j = 0; j < KB2; ++j) f(ITEM_TS, k, j, part);
    if (k + 2 < KS)
      for (int j = 0; j < kb1; ++j) f(ITEM_P, k + 2, j, 0);
  }
}

template <int HS, int O>
struct CfgH {
  static_assert(HS == 64 || HS == 128, "hidden width of the fused AGCN kernel: 64 or 128");
  static_assert(O == HS || O == 2 * HS, "output width: HS (update) or 2*HS (gate)");
  static constexpr uint32_t A_SLOT = BM * 128;                    // 16 KB: [128 rows][64 halves]
  static constexpr uint32_t B_SLOT = (uint32_t)O * 128;           // [O rows][64 halves]  (>= the [HS rows] tile of MMA1)
  static constexpr uint32_t STAGE = A_SLOT + B_SLOT;
  static constexpr int NST = O >= 256 ? 4 : (O >= 128 ? 6 : 8);
  static constexpr uint32_t SCRATCH = 4 * 32 * 36 * 4;            // rounding warps: 32 x 36 floats each
  static constexpr size_t SMEM = (size_t)NST * STAGE + SCRATCH + 1024;
  static_assert((size_t)NST * STAGE >= EPI_WARPS * 32 * 36 * 4, "the epilogue stages through the (idle) ring");
  static constexpr uint32_t TM_ACC = 0, TM_P0 = O, TM_P1 = O + HS;
  static constexpr uint32_t TMEM_COLS = pow2_cols(O + 2 * HS);
  static constexpr int KB2 = HS / BKH;                            // k-blocks of the weight contraction per segment
};

// ---- epilogue functors ------------------------------------------------------------------------
// load4 / fin4 work on 4 consecutive columns of one (node, b) row; fin4 also writes the fp16 operand copy the next fused
// launch consumes (row-major only: the propagation reads it as an MN-major operand).

// Gate AGCN (model/MegaCRN.py:43-45): zr = sigmoid(acc); columns [0,H) = z, [H,2H) = r; writes z, r, z*h.
struct EpiGateH {
  static constexpr int NP = 1;
  int H;
  const float* h;       // [R][H] exact state
  float* z;             // null in eval
  float* r;
  float* zh32;          // fp32 copy of the fp16-rounded z*h (XPu block 0, read by the backward); null in eval
  __half* x16;          // [R][H]  z*h: A operand of the update AGCN's identity segment and (MN-major) B operand of its propagation
  __device__ __forceinline__ void load4(int row, int n0, float4 (&p)[NP]) const {
    p[0] = (n0 < H) ? ldg4(h + (int64_t)row * H + n0) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ void fin4(int row, int n0, const float4 (&p)[NP], const float (&acc)[4]) const {
    float st[4];
    float s0, s1, s2, s3;
    s0 = sigmoid_fast(acc[0]); s1 = sigmoid_fast(acc[1]); s2 = sigmoid_fast(acc[2]); s3 = sigmoid_fast(acc[3]);
    if (n0 < H) {
      const int64_t o = (int64_t)row * H + n0;
      if (z) st4(z + o, s0, s1, s2, s3);
      st[0] = round_h(s0 * p[0].x); st[1] = round_h(s1 * p[0].y); st[2] = round_h(s2 * p[0].z); st[3] = round_h(s3 * p[0].w);
      if (zh32) st4(zh32 + o, st[0], st[1], st[2], st[3]);
      *reinterpret_cast<uint2*>(x16 + o) = make_uint2(pack_h2(st[0], st[1]), pack_h2(st[2], st[3]));
    } else {
      st4(r + (int64_t)row * H + (n0 - H), s0, s1, s2, s3);
    }
  }
};

// Update AGCN (model/MegaCRN.py:46-47): hc = tanh(acc); h' = r*h + (1-r)*hc.
struct EpiUpdateH {
  static constexpr int NP = 2;
  int H;
  const float* h; const float* r;
  float* hc;            // null in eval
  float* h_out;         // exact new state
  float* h32;           // fp32 copy of the fp16-rounded new state (next step's XPg block 0); null in eval / last step
  __half* x16;          // next step's operand copy [R][H]; null after the last step
  __device__ __forceinline__ void load4(int row, int n0, float4 (&p)[NP]) const {
    const int64_t o = (int64_t)row * H + n0;
    p[0] = ldg4(h + o);
    p[1] = ldg4(r + o);
  }
  __device__ __forceinline__ void fin4(int row, int n0, const float4 (&p)[NP], const float (&acc)[4]) const {
    float st[4];
    const int64_t o = (int64_t)row * H + n0;
    const float4 hv = p[0], rv = p[1];
    const float c0 = tanh_fast(acc[0]), c1 = tanh_fast(acc[1]), c2 = tanh_fast(acc[2]), c3 = tanh_fast(acc[3]);
    if (hc) st4(hc + o, c0, c1, c2, c3);
    const float n0_ = rv.x * hv.x + (1.0f - rv.x) * c0, n1_ = rv.y * hv.y + (1.0f - rv.y) * c1;
    const float n2_ = rv.z * hv.z + (1.0f - rv.z) * c2, n3_ = rv.w * hv.w + (1.0f - rv.w) * c3;
    st4(h_out + o, n0_, n1_, n2_, n3_);
    st[0] = round_h(n0_); st[1] = round_h(n1_); st[2] = round_h(n2_); st[3] = round_h(n3_);
    if (h32) st4(h32 + o, st[0], st[1], st[2], st[3]);
    if (x16) *reinterpret_cast<uint2*>(x16 + o) = make_uint2(pack_h2(st[0], st[1]), pack_h2(st[2], st[3]));
  }
};

#define MCRN_TLH(slot)                                                     \
  do {                                                                     \
    if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0) p.dbg[(slot)] = clock64(); \
  } while (0)

template <int HS, int O, class Epi>
__global__ void __launch_bounds__(FTHREADS, 1)
agcn_fused_h_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmXB,
                    const __grid_constant__ CUtensorMap tmXA, const __grid_constant__ CUtensorMap tmIB,
                    const __grid_constant__ CUtensorMap tmW, HParams p, Epi epi) {
  using C = CfgH<HS, O>;
  constexpr int NST = C::NST;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[NST];
  __shared__ __align__(8) uint64_t empty_bar[NST];
  __shared__ __align__(8) uint64_t p_full_bar[2];    // MMA1 of a P buffer retired (tcgen05.commit)
  __shared__ __align__(8) uint64_t p_ready_bar[2];   // the 4 rounding warps have packed the P buffer
  __shared__ __align__(8) uint64_t acc_full_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, b = blockIdx.y;
  const int kb1 = (p.N + BKH - 1) / BKH;
  const int NSEG = p.KS + 2;                         // weight segments per part
  if (threadIdx.x == 0) MCRN_TLH(0);
  if (p.span != nullptr && threadIdx.x == 0) atomicMin(p.span, fused::globaltimer_ns());
  if (p.dbg != nullptr && threadIdx.x == 0) {        // per-CTA wall-clock stamps (ns): [512 + 2 cta] = start, [513 + 2 cta] = end
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.dbg[512 + 2 * (blockIdx.y * gridDim.x + blockIdx.x)] = (long long)gt;
  }

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmS) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmXB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmXA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmIB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
#pragma unroll
    for (int s = 0; s < NST; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&p_full_bar[0]), 1);
    mbar_init(smem_u32(&p_full_bar[1]), 1);
    mbar_init(smem_u32(&p_ready_bar[0]), 4);
    mbar_init(smem_u32(&p_ready_bar[1]), 4);
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
  if (threadIdx.x == 0) MCRN_TLH(1);
  // programmatic dependent launch: everything above overlapped the previous kernel's tail; its results are read from here on
  pdl_wait();
  if (!p.pdl_late) pdl_launch_dependents();

  const int pw = warp == 0 ? 0 : (warp >= 2 + EPI_WARPS ? warp - (2 + EPI_WARPS) + 1 : -1);   // producer index or -1
  if (pw >= 0) {
    if (lane == 0) {                                     // ===== TMA producers: item `it` belongs to producer it % NPROD =====
      int it = 0;
      for_each_item_h<C::KB2>(p.KS, kb1, p.nparts, p.ib_blocks, [&](int type, int k, int j, int part) {
        if (it % NPROD != pw) { ++it; return; }
        const int s = it % NST;
        if (it >= NST) mbar_wait_b(smem_u32(&empty_bar[s]), (((uint32_t)(it / NST)) & 1u) ^ 1u);
        if (it < 200) MCRN_TLH(240 + it);
        const uint32_t fb = smem_u32(&full_bar[s]);
        const uint32_t a_dst = smem_base + (uint32_t)s * C::STAGE, b_dst = a_dst + C::A_SLOT;
        if (type == ITEM_P) {
          mbar_expect_tx(fb, C::A_SLOT + (uint32_t)HS * 128);
          tma_load_4d(a_dst, &tmS, fb, j * BKH, m0, k, 0);                      // S_k[m0.., 64 j..]
#pragma unroll
          for (int q = 0; q < HS / 64; ++q)                                     // X rows (64 j.., b), channels 64 q..  (MN-major B)
            tma_load_4d(b_dst + q * 8192, &tmXB, fb, q * 64, b, j * BKH, 0);
        } else {
          const int wseg = (type == ITEM_SS ? k : 1 + k) + part * NSEG;
          if (type == ITEM_SS) {
            mbar_expect_tx(fb, C::A_SLOT + C::B_SLOT);
            tma_load_4d(a_dst, k == 0 ? &tmXA : &tmIB, fb, j * BKH, b, m0, 0);  // X / IB rows (m0.., b), channels 64 j..
          } else {
            mbar_expect_tx(fb, C::B_SLOT);
          }
          tma_load_4d(b_dst, &tmW, fb, j * BKH, 0, wseg, 0);                    // W^T[wseg][0..O][64 j..]
        }
        ++it;
      });
      if (pw == 0) MCRN_TLH(232);
    }
  } else if (warp == 1) {
    if (lane == 0) {                                     // ===== MMA issuer =====
      constexpr uint32_t idesc1 = make_idesc_f16<HS>() | (1u << 16);     // B of MMA1 is MN-major (row-major X rows)
      constexpr uint32_t idesc2 = make_idesc_f16<O>();
      int it = 0;
      bool acc_on = false;
      for_each_item_h<C::KB2>(p.KS, kb1, p.nparts, p.ib_blocks, [&](int type, int k, int j, int part) {
        const int s = it % NST;
        mbar_wait_b(smem_u32(&full_bar[s]), ((uint32_t)(it / NST)) & 1u);
        tcgen05_fence_after();
        if (it < 200) MCRN_TLH(2 + it);
        const uint32_t a_addr = smem_base + (uint32_t)s * C::STAGE, b_addr = a_addr + C::A_SLOT;
        const uint32_t pbuf = tmem_base + ((k & 1) ? C::TM_P1 : C::TM_P0);
        if (type == ITEM_P) {
          // UMMA_K = 16 for fp16: 32 bytes along the swizzled row; the last k-block stops at the node count (the tensor
          // maps zero-fill beyond N, so whole 16-node steps past it would only add zeros)
          const int nkk = min(BKH / 16, (p.N - j * BKH + 15) / 16);
#pragma unroll
          for (int kk = 0; kk < BKH / 16; ++kk) {
            if (kk < nkk) {
              const uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024, 2);
              const uint64_t bd = make_smem_desc(b_addr + kk * 2048, 8192, 1024, 2);     // 16 node rows per step, 64-channel blocks 8 KB apart
              tcgen05_mma_f16(pbuf, ad, bd, idesc1, (j > 0 || kk > 0) ? 1u : 0u);
            }
          }
          tcgen05_commit(smem_u32(&empty_bar[s]));
          if (j == kb1 - 1) tcgen05_commit(smem_u32(&p_full_bar[k & 1]));
        } else if (type == ITEM_SS) {
#pragma unroll
          for (int kk = 0; kk < BKH / 16; ++kk) {
            const uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024, 2);
            const uint64_t bd = make_smem_desc(b_addr + kk * 32, 16, 1024, 2);
            tcgen05_mma_f16(tmem_base + C::TM_ACC, ad, bd, idesc2, (acc_on || kk > 0) ? 1u : 0u);
          }
          acc_on = true;
          tcgen05_commit(smem_u32(&empty_bar[s]));
        } else {
          if (j == 0 && part == 0) {                     // P_k has been packed to fp16 in place by the rounding warps
            mbar_wait_b(smem_u32(&p_ready_bar[k & 1]), ((uint32_t)(k >> 1)) & 1u);
            tcgen05_fence_after();
          }
#pragma unroll
          for (int kk = 0; kk < BKH / 16; ++kk) {        // 16 halves of K = 8 packed TMEM columns
            const uint64_t bd = make_smem_desc(b_addr + kk * 32, 16, 1024, 2);
            tcgen05_mma_f16_ts(tmem_base + C::TM_ACC, pbuf + (uint32_t)(j * (BKH / 2) + kk * 8), bd, idesc2, (acc_on || kk > 0) ? 1u : 0u);
          }
          acc_on = true;
          tcgen05_commit(smem_u32(&empty_bar[s]));
        }
        ++it;
      });
      tcgen05_commit(smem_u32(&acc_full_bar));
      MCRN_TLH(233);
    }
  } else if (warp < 2 + EPI_WARPS) {                     // ===== rounding + epilogue warps =====
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may access
    const int ew = warp - 2, half_id = ew >> 2;          // ew 0..15; warps 2..5 (half_id 0) also do the P rounding
    const int cq = (lane & 7) * 4, r0 = lane >> 3;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int node0 = m0 + quarter * 32;
    if (half_id == 0) {
      // ---- P_k: fp32 accumulator -> fp16 pairs packed in place; training: store the rounded block for the backward ----
      float* scr = reinterpret_cast<float*>(smem_al + (size_t)NST * C::STAGE) + ew * (32 * 36);
      for (int k = 0; k < p.KS; ++k) {
        mbar_wait_b(smem_u32(&p_full_bar[k & 1]), ((uint32_t)(k >> 1)) & 1u);
        tcgen05_fence_after();
        if (warp == 2 && lane == 0 && k < 5) MCRN_TLH(210 + 4 * k);
        const uint32_t pbuf = tmem_base + ((k & 1) ? C::TM_P1 : C::TM_P0) + lane_off;
#pragma unroll 1
        for (int c = 0; c < HS / 32; ++c) {
          float v[32];
          tmem_ld_32x32b_x32(pbuf + (uint32_t)(c * 32), v);
          uint32_t u[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) u[i] = pack_h2(v[2 * i], v[2 * i + 1]);
          tmem_st_32x32b_x16(pbuf + (uint32_t)(c * 16), u);      // columns [16c, 16c+16) <= columns already read
          if (p.xp_save != nullptr && node0 < p.N) {
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<float4*>(&scr[lane * 36 + i]) = make_float4(round_h(v[i]), round_h(v[i + 1]), round_h(v[i + 2]), round_h(v[i + 3]));
            __syncwarp();
            float* dst = p.xp_save + (int64_t)(1 + k) * p.blk_stride + c * 32 + cq;
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
        if (lane == 0) mbar_arrive(smem_u32(&p_ready_bar[k & 1]));
        if (warp == 2 && lane == 0 && k < 5) MCRN_TLH(211 + 4 * k);
      }
    }
    // ---- epilogue: accumulator -> gate / update math; rows = (node, b); this warp takes chunks c = half_id (mod 4) ----
    mbar_wait_b(smem_u32(&acc_full_bar), 0);
    tcgen05_fence_after();
    if (p.pdl_late) pdl_launch_dependents();
    if (warp == 2 && lane == 0) MCRN_TLH(230);
    if (node0 < p.N) {
      float* scr = reinterpret_cast<float*>(smem_al) + ew * (32 * 36);      // the ring is idle now
#pragma unroll 1
      for (int c = half_id; c < O / 32; c += EPI_WARPS / 4) {
        float v[32];
        tmem_ld_32x32b_x32(tmem_base + C::TM_ACC + lane_off + (uint32_t)(c * 32), v);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(&scr[lane * 36 + i]) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        __syncwarp();
        const int col = c * 32 + cq;
        constexpr int RB = Epi::NP <= 1 ? 8 : 4;
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
        __syncwarp();                                    // the staging tile is reused by the next chunk
      }
    }
  }
  if (warp == 2 && lane == 0) MCRN_TLH(231);
  tcgen05_fence_before();
  __syncthreads();
  if (p.dbg != nullptr && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    p.dbg[513 + 2 * (blockIdx.y * gridDim.x + blockIdx.x)] = (long long)gt;
  }
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
  if (p.span != nullptr && threadIdx.x == 0) atomicMax(p.span + 1, fused::globaltimer_ns());
}

// ---- operand conversion kernels ---------------------------------------------------------------
// supports fp32 [KS][N][ld] -> fp16 [KS][N][ld16]
__global__ void k_supports_to_half(const float* __restrict__ S, __half* __restrict__ S16, int rows, int n, int ld, int ld16) {
  const int64_t total = (int64_t)rows * ld16;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld16;
    const int c = (int)(i - r * ld16);
    S16[i] = __float2half_rn(c < n ? S[r * ld + c] : 0.f);
  }
}
// state fp32 [N][B][HS] -> fp16 row-major [N][B][HS]
__global__ void k_state_to_half(const float* __restrict__ x, __half* __restrict__ x16, int64_t total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    x16[i] = __float2half_rn(x[i]);
}
// folded weights fp32 [2 (TF32 hi, lo)][KS+2][HS][O] -> fp16 hi / lo, transposed: [2][KS+2][O][HS]
__global__ void k_weights_to_half(const float* __restrict__ wall, __half* __restrict__ w16, int nseg, int HS, int O) {
  const int64_t total = (int64_t)nseg * HS * O;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % HS);
    const int o = (int)((i / HS) % O);
    const int seg = (int)(i / ((int64_t)HS * O));
    const int64_t src = ((int64_t)seg * HS + c) * O + o;
    const float w = wall[src] + wall[total + src];            // hi + lo of the TF32 split = the weight to ~21 bits
    const __half hi = __float2half_rn(w);
    w16[i] = hi;
    w16[total + i] = __float2half_rn(w - __half2float(hi));
  }
}

// ---- compact input block ----------------------------------------------------------------------
// The "input block" of an AGCN contraction (input channels of every support + the bias one) has only NB*Cin + 1 <= 16
// non-zero columns.  Compact form: ib16c [R][64] halves (one 128-byte swizzle row per (node, b): exactly one k-block of
// the fused kernel) and, for the weight gradient of that block, ib32c [R][16] floats.
constexpr int IBC = 64, IBF = 16;

// Encoder, all steps at once.  xpin: [NB][N][T][B][Cin] (block 0 = staged input, 1.. = propagated).
// steps [t0, t0 + nt): step 0 is built on the main stream, the others beside the first encoder cell
__global__ void k_encoder_input_blocks(const float* __restrict__ xpin, int NB, int N, int T, int B, int Cin,
                                       __half* __restrict__ ib16c, float* __restrict__ ib32c, int t0, int nt) {
  const int64_t R = (int64_t)N * B, total = (int64_t)nt * R * IBC;
  const int nin = NB * Cin;
  ib16c += (int64_t)t0 * R * IBC;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % IBC);
    const int64_t row = (i / IBC) % R;
    const int t = t0 + (int)(i / ((int64_t)IBC * R));
    float v = 0.f;
    if (j < nin) {
      const int k = j / Cin, ci = j - k * Cin;
      const int n = (int)(row / B), b = (int)(row - (int64_t)n * B);
      v = round_h(xpin[((((int64_t)k * N + n) * T + t) * B + b) * Cin + ci]);
    } else if (j == nin) {
      v = 1.0f;
    }
    ib16c[i] = __float2half_rn(v);
    if (ib32c != nullptr && j < IBF) ib32c[((int64_t)t * R + row) * IBF + j] = v;
  }
}

// Decoder step t: staging of [go | y_cov[:, t]] (model/MegaCRN.py:185), its propagation through the KS supports and the
// compact input block, in one kernel.  go_src: nullptr (zeros) or a [B][T][N][Cout] tensor read at step t-1.
// S: exact fp32 supports [KS][N][ldS].  A block owns DI_NODES nodes x (DI_COLS / Cd) batch elements: it stages the
// DI_COLS input columns (all N source nodes) and the KS * DI_NODES support rows in shared memory with coalesced loads,
// accumulates, assembles the complete 128-byte rows of the input block in shared memory and writes them with 16-byte
// stores.  Needs KS * DI_NODES <= 8 * DI_RPT rows (KS <= 4 with the constants below) and (KS + 1) * Cd + 1 <= 16.
constexpr int DI_COLS = 32, DI_NODES = 6, DI_RPT = 3, DI_ROWS = 8 * DI_RPT;
__global__ void __launch_bounds__(256) k_decoder_input_block(const float* __restrict__ go_src, const float* __restrict__ ycov,
                                                             const float* __restrict__ S, int ldS, int KS, int N, int B, int T,
                                                             int Cout, int Ycov, unsigned tmask, float* __restrict__ xin_base,
                                                             int64_t xin_step, __half* __restrict__ ib16_base,
                                                             float* __restrict__ ib32_base) {
  // blockIdx.z selects the z-th set bit of tmask = the decoder step this block builds (all teacher-forced steps of a
  // forward are built by ONE launch; a free-running step by a launch of its own once the previous output exists)
  int t = 0;
  {
    unsigned m = tmask;
    for (int z = blockIdx.z; z > 0; --z) m &= m - 1;
    t = __ffs(m) - 1;
  }
  const int64_t Rr = (int64_t)N * B;
  float* xin_out = xin_base ? xin_base + (int64_t)t * xin_step : nullptr;
  __half* ib16c = ib16_base + (int64_t)t * Rr * IBC;
  float* ib32c = ib32_base ? ib32_base + (int64_t)t * Rr * IBF : nullptr;
  if (t == 0) go_src = nullptr;                        // go_0 = 0 (model/MegaCRN.py:182)
  extern __shared__ float sh[];                        // xs [DI_COLS][N|1], ss [DI_ROWS][N + 1], os [DI_NODES * DI_COLS][IBF + 1]
  const int Cd = Cout + Ycov, cols = B * Cd, nin = (KS + 1) * Cd;
  const int c0 = blockIdx.x * DI_COLS, n0 = blockIdx.y * DI_NODES;
  const int xld = N | 1;                               // odd leading dimension: column-major input tile without bank conflicts
  float* xs = sh;                                      // xs[cl * xld + m]
  float* ss = xs + xld * DI_COLS;
  float* os = ss + DI_ROWS * (N + 1);
  auto xin = [&](int m, int col) -> float {
    const int b = col / Cd, c = col - b * Cd;
    if (c < Cout) return go_src ? __ldg(go_src + (((int64_t)b * T + (t - 1)) * N + m) * Cout + c) : 0.f;
    return __ldg(ycov + (((int64_t)b * T + t) * N + m) * Ycov + (c - Cout));
  };
  // staging: threads along the source node m (coalesced); per input column / support row the base address is
  // block-uniform (no per-element index arithmetic); 8 independent loads are issued before their shared-memory stores --
  // the block is alone on its SM and bound by load latency, not bandwidth
  for (int mb = 0; mb < N; mb += blockDim.x) {
    const int m = mb + threadIdx.x;
    for (int cb = 0; cb < DI_COLS; cb += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int col = c0 + cb + u;
        const int b = col / Cd, c = col - b * Cd;
        v[u] = 0.f;
        if (m < N && col < cols) {
          if (c < Cout) { if (go_src) v[u] = __ldg(go_src + (((int64_t)b * T + (t - 1)) * N + m) * Cout + c); }
          else v[u] = __ldg(ycov + (((int64_t)b * T + t) * N + m) * Ycov + (c - Cout));
        }
      }
      if (m < N) {
#pragma unroll
        for (int u = 0; u < 8; ++u) xs[(cb + u) * xld + m] = v[u];
      }
    }
    for (int rb = 0; rb < DI_ROWS; rb += 8) {                       // local row rl = k * DI_NODES + nl
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int rl = rb + u, k = rl / DI_NODES, n = n0 + (rl - k * DI_NODES);
        v[u] = (m < N && k < KS && n < N) ? __ldg(S + ((int64_t)k * N + n) * ldS + m) : 0.f;
      }
      if (m < N) {
#pragma unroll
        for (int u = 0; u < 8; ++u) ss[(rb + u) * (N + 1) + m] = v[u];
      }
    }
  }
  for (int i = threadIdx.x; i < DI_NODES * DI_COLS * (IBF + 1); i += blockDim.x) os[i] = 0.f;
  __syncthreads();
  const int cl = threadIdx.x & (DI_COLS - 1), rg = threadIdx.x / DI_COLS;     // 8 row lanes x 32 columns
  float acc[DI_RPT];
#pragma unroll
  for (int i = 0; i < DI_RPT; ++i) acc[i] = 0.f;
  for (int m = 0; m < N; ++m) {
    const float xv = xs[cl * xld + m];
#pragma unroll
    for (int i = 0; i < DI_RPT; ++i) acc[i] = fmaf(ss[(rg * DI_RPT + i) * (N + 1) + m], xv, acc[i]);
  }
  // stage the block's values: local output row = nl * (DI_COLS / Cd) + b_local, column j of the input block
  const int bl = cl / Cd, c = cl - bl * Cd, bpb = DI_COLS / Cd;
#pragma unroll
  for (int i = 0; i < DI_RPT; ++i) {
    const int rl = rg * DI_RPT + i, k = rl / DI_NODES, nl = rl - k * DI_NODES;
    if (k < KS) os[(nl * bpb + bl) * (IBF + 1) + (1 + k) * Cd + c] = round_h(acc[i]);
  }
  for (int i = threadIdx.x; i < DI_NODES * DI_COLS; i += blockDim.x) {        // block 0 = the input itself, bias one
    const int nl = i / DI_COLS, cc = i - nl * DI_COLS, b2 = cc / Cd, c2 = cc - b2 * Cd;
    if (n0 + nl < N) {
      os[(nl * bpb + b2) * (IBF + 1) + c2] = xs[cc * xld + n0 + nl];           // raw; rounded at the store below
      if (c2 == 0) os[(nl * bpb + b2) * (IBF + 1) + nin] = 1.0f;
    }
  }
  __syncthreads();
  // write complete rows: ib16c row = 8 x 16 bytes (chunks 0,1 from the staged values, the rest zero), ib32c row = 4 x float4
  const int b0 = c0 / Cd;
  for (int i = threadIdx.x; i < DI_NODES * bpb * 8; i += blockDim.x) {
    const int q = i & 7, lr = i >> 3, nl = lr / bpb, b2 = lr - nl * bpb;
    const int n = n0 + nl, b = b0 + b2;
    if (n >= N || b >= B) continue;
    const int64_t row = (int64_t)n * B + b;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (q < 2) {
      const float* o = os + lr * (IBF + 1) + q * 8;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (q * 8 + e < Cd) ? round_h(o[e]) : o[e];
      u = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
      if (ib32c != nullptr) {
        st4(ib32c + row * IBF + q * 8, v[0], v[1], v[2], v[3]);
        st4(ib32c + row * IBF + q * 8 + 4, v[4], v[5], v[6], v[7]);
      }
      if (xin_out != nullptr && q == 0)
        for (int e = 0; e < Cd; ++e) xin_out[row * Cd + e] = tf32_rn(o[e]);   // operand of the TF32 dS / dxin GEMMs of the backward
    }
    *reinterpret_cast<uint4*>(ib16c + row * IBC + q * 8) = u;
  }
}

// ---- host side ----------------------------------------------------------------------------------
int encode_tensor_map_h(CUtensorMap* out, const void* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                        const uint32_t box[4]);

static inline int ld_half(int n) { return (n + 7) / 8 * 8; }     // 16-byte row stride for fp16 rows

struct HOperands {
  const __half* S16;      // [KS][N][ld_half(N)]
  const __half* X16;      // [R][HS]
  const __half* IB16;     // [R][ib_ld]: ib_ld = HS (full input block) or 64 (compact: only the first k-block is non-zero)
  const __half* W16;      // [nparts][KS+2][O][HS]
  float* xp_save;         // fp32 XP buffer (training) or null
  int ib_ld = 0;          // 0 = HS
};

template <int HS, int O, class Epi>
int launch_agcn_fused_h(int N, int B, int KS, const HOperands& op, int nparts, const Epi& epi, cudaStream_t st) {
  using C = CfgH<HS, O>;
  const int64_t R = (int64_t)N * B;
  const int ldn = ld_half(N);
  CUtensorMap tS, tXB, tXA, tIB, tW;
  {
    uint64_t dims[4] = {(uint64_t)N, (uint64_t)N, (uint64_t)KS, 1};
    uint64_t str[3] = {(uint64_t)ldn * 2, (uint64_t)N * ldn * 2, (uint64_t)KS * N * ldn * 2};
    uint32_t box[4] = {BKH, BM, 1, 1};
    MCRN_TRY(encode_tensor_map_h(&tS, op.S16, dims, str, box));
  }
  {
    uint64_t dims[4] = {(uint64_t)HS, (uint64_t)B, (uint64_t)N, 1};
    uint64_t str[3] = {(uint64_t)HS * 2, (uint64_t)B * HS * 2, (uint64_t)R * HS * 2};
    uint32_t box[4] = {BKH, 1, BM, 1};
    MCRN_TRY(encode_tensor_map_h(&tXA, op.X16, dims, str, box));
    uint32_t boxb[4] = {64, 1, BKH, 1};                  // MN-major B of MMA1: 64 channels x 64 node rows of batch element b
    MCRN_TRY(encode_tensor_map_h(&tXB, op.X16, dims, str, boxb));
    const uint64_t il = op.ib_ld ? (uint64_t)op.ib_ld : (uint64_t)HS;
    uint64_t dimi[4] = {il, (uint64_t)B, (uint64_t)N, 1};
    uint64_t stri[3] = {il * 2, (uint64_t)B * il * 2, (uint64_t)R * il * 2};
    MCRN_TRY(encode_tensor_map_h(&tIB, op.IB16, dimi, stri, box));
  }
  {
    uint64_t dims[4] = {(uint64_t)HS, (uint64_t)O, (uint64_t)(nparts * (KS + 2)), 1};
    uint64_t str[3] = {(uint64_t)HS * 2, (uint64_t)O * HS * 2, (uint64_t)nparts * (KS + 2) * O * HS * 2};
    uint32_t box[4] = {BKH, (uint32_t)O, 1, 1};
    MCRN_TRY(encode_tensor_map_h(&tW, op.W16, dims, str, box));
  }
  HParams p;
  p.N = N; p.B = B; p.KS = KS; p.nparts = nparts;
  p.ib_blocks = (op.ib_ld ? op.ib_ld : HS) / BKH;
  p.xp_save = op.xp_save;
  p.blk_stride = R * HS;
  p.dbg = nullptr;
  p.pdl_late = (g_pdl_chain >> 3) & 1;
  p.span = fused::next_span();
  if (fused::g_dbg_timeline != nullptr) {
    if (fused::g_dbg_which < 0 || fused::g_dbg_count == fused::g_dbg_which) p.dbg = fused::g_dbg_timeline;
    ++fused::g_dbg_count;
  }
  auto kern = agcn_fused_h_kernel<HS, O, Epi>;
  static bool attr_set = false;
  if (!attr_set) {
    MCRN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    attr_set = true;
  }
  dim3 grid(ceil_div(N, BM), B, 1);
  const int pi = fused::prof_begin(fused::prof_class(0, HS, O == HS ? 1 : 0), st);
  MCRN_TRY(launch_chain(1, kern, grid, dim3(FTHREADS), C::SMEM, st, "agcn_fused_h_kernel", tS, tXB, tXA, tIB, tW, p, epi));
  fused::prof_end(pi, st);
  return MCRN_OK;
}

}  // namespace fusedh
}  // namespace mcrn
