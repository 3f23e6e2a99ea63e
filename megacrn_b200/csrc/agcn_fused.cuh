// Fused AGCN forward for sm_100a: the K-hop graph convolution (model/MegaCRN.py:24-26) and the weight contraction
// (:27) of ONE AGCN call, with the gate / update elementwise tail of the AGCRN cell (:43-47) in the epilogue, in a
// single kernel.  The propagated blocks P_k = S_k * X never leave the SM:
//
//   CTA (node tile of 128 nodes, batch element b):
//     MMA1  P_k[128 x HS]  = S_k[tile rows, :] * X[:, b, :]          A = support rows (TMA, K-major, smem)
//                                                                     B = the state of batch element b (TMA, smem)
//           accumulator in TMEM (two P buffers, ping-pong over k)
//     round P_k -> TF32 (round-to-nearest) in place in TMEM (tcgen05.ld / cvt.rna / tcgen05.st) by the epilogue
//           warps; in training mode the rounded block is also stored for the backward (XP block 1+k)
//     MMA2  acc[128 x O]  += P_k * W_k            A = P_k straight from TMEM (tcgen05.mma, A in tensor memory)
//                          + X_tile * W_0 + IB_tile * W_NB      (identity block and input-channel/bias block: A in smem)
//           B = folded weights (TMA, smem), TF32 hi (+ lo residual) parts
//     epilogue  acc -> sigmoid / tanh, z*h or the GRU blend (EpiGate / EpiUpdate of gemm.cuh)
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = rounding + epilogue.
// One ring of shared-memory stages feeds both MMA phases; every ring item is "A slot (16 KB) + B slot (O*128 B)".
#pragma once

#include "gemm_tc.cuh"

namespace mcrn {
namespace fused {

using namespace tc;

struct FusedParams {
  int N, B, KS;        // nodes, batch, number of real supports
  int nparts;          // 1 = TF32 hi weights only, 2 = hi + lo residual
  float* xp_save;      // training: base of the XP buffer [NB+1][R][HS]; P_k is stored to block 1+k.  null = eval
  int64_t blk_stride;  // R * HS
  long long* dbg;      // debug: clock64 timestamps of CTA (0,0) (mcrn_debug_fused_timeline); null in production
};
extern long long* g_dbg_timeline;   // host side: non-null only while mcrn_debug_fused_timeline is armed
extern int g_dbg_which, g_dbg_count;
extern unsigned long long* g_dbg_span;   // per-launch wall-clock spans {min CTA start, max CTA end} (ns, %globaltimer), mcrn_debug_launch_spans
extern int g_dbg_span_n, g_dbg_span_cap;
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
static inline unsigned long long* next_span() { return (g_dbg_span != nullptr && g_dbg_span_n < g_dbg_span_cap) ? g_dbg_span + 2 * (g_dbg_span_n++) : nullptr; }

// timeline slots: [0] start, [1] after prologue, [2 + it] MMA issuer: operands of item `it` landed (up to 200 items),
// [210 + 4k .. ] rounding warp 2: p_full seen / rounded+stored / (2 unused), [230] acc_full seen, [231] epilogue done,
// [232] producer done issuing, [233] MMA issuer done issuing, [240 + it] producer: slot free for item `it`
#define MCRN_TL(slot)                                                      \
  do {                                                                     \
    if (p.dbg != nullptr && blockIdx.x == 0 && blockIdx.y == 0) p.dbg[(slot)] = clock64(); \
  } while (0)

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must fail loudly (trap -> launch error), never hang the device.
__device__ __forceinline__ void mbar_wait_b(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000ll) {
      printf("agcn_fused: mbarrier timeout (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x,
             bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tcgen05_mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const float (&v)[32]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

enum : int { ITEM_P = 0, ITEM_SS = 1, ITEM_TS = 2 };

// ---- per-kernel-class timing (bench.py roofline): CUDA events recorded on the launching stream around every fused
// launch while enabled (eager launches only; a capturing stream is left alone).  Classes: HS/O/direction, see prof_class.
struct KernelProf {
  static constexpr int NCLS = 16, NEV = 1024;
  int enabled = 0;
  int count = 0;
  cudaEvent_t ev[NEV][2];
  int cls[NEV];
  bool created = false;
};
extern KernelProf g_prof;
// class id: direction (0 fwd, 1 bwd) * 8 + (HS == 128 ? 4 : 0) + variant (fwd: 0 gate / 1 update; bwd: 0 BU / 1 BG)
static inline int prof_class(int bwd, int HS, int variant) { return bwd * 8 + (HS == 128 ? 4 : 0) + variant; }
static inline int prof_begin(int cls, cudaStream_t st) {
  if (!g_prof.enabled || g_prof.count >= KernelProf::NEV) return -1;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return -1;
  if (!g_prof.created) {
    for (int i = 0; i < KernelProf::NEV; ++i) { cudaEventCreate(&g_prof.ev[i][0]); cudaEventCreate(&g_prof.ev[i][1]); }
    g_prof.created = true;
  }
  const int i = g_prof.count++;
  g_prof.cls[i] = cls;
  cudaEventRecord(g_prof.ev[i][0], st);
  return i;
}
static inline void prof_end(int i, cudaStream_t st) { if (i >= 0) cudaEventRecord(g_prof.ev[i][1], st); }

constexpr int pow2_cols(int c) { return c <= 32 ? 32 : c <= 64 ? 64 : c <= 128 ? 128 : c <= 256 ? 256 : 512; }

template <int HS, int O>
struct Cfg {
  static_assert(HS == 64 || HS == 128, "hidden width of the fused AGCN kernel: 64 or 128");
  static_assert(O == HS || O == 2 * HS, "output width: HS (update) or 2*HS (gate)");
  static constexpr uint32_t A_SLOT = BM * BK * 4;                 // 16 KB: [128 rows][32 k] fp32, 128B-swizzled
  static constexpr uint32_t B_SLOT = (uint32_t)O * BK * 4;        // O/32 slabs of [32 k][32 n] (>= the HS/32 slabs of MMA1)
  static constexpr uint32_t STAGE = A_SLOT + B_SLOT;
  static constexpr int NST = O >= 256 ? 4 : (O >= 128 ? 5 : 6);
  static constexpr uint32_t SCRATCH = 4 * 32 * 36 * 4;            // per epilogue warp: 32 x 36 floats (transposition)
  static constexpr size_t SMEM = (size_t)NST * STAGE + SCRATCH + 1024;
  static constexpr uint32_t TM_ACC = 0, TM_P0 = O, TM_P1 = O + HS;
  static constexpr uint32_t TMEM_COLS = pow2_cols(O + 2 * HS);
  static constexpr int KB2 = HS / BK;                             // k-blocks of the weight contraction per segment
};

// The ring items of one CTA, in issue order (identical in the producer and in the MMA issuer).
//   P(k)      : kb1 items   (support k-block j, state k-block j)
//   SS(blk)   : nparts * KB2 items (A = XP block `blk` tile from smem, B = weight segment)
//   TS(k)     : nparts * KB2 items (A = P_k from TMEM, B = weight segment 1+k)
template <int KB2, class F>
__device__ __forceinline__ void for_each_item(int KS, int kb1, int nparts, F&& f) {
  for (int j = 0; j < kb1; ++j) f(ITEM_P, 0, j, 0);
  if (KS > 1)
    for (int j = 0; j < kb1; ++j) f(ITEM_P, 1, j, 0);
  for (int part = 0; part < nparts; ++part)
    for (int j = 0; j < KB2; ++j) f(ITEM_SS, 0, j, part);
  for (int part = 0; part < nparts; ++part)
    for (int j = 0; j < KB2; ++j) f(ITEM_SS, KS + 1, j, part);
  for (int k = 0; k < KS; ++k) {
    for (int part = 0; part < nparts; ++part)
      for (int j = 0; j < KB2; ++j) f(ITEM_TS, k, j, part);
    if (k + 2 < KS)
      for (int j = 0; j < kb1; ++j) f(ITEM_P, k + 2, j, 0);
  }
}

template <int HS, int O, class Epi>
__global__ void __launch_bounds__(THREADS, 1)
agcn_fused_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmXA,
                  const __grid_constant__ CUtensorMap tmXB, const __grid_constant__ CUtensorMap tmW, FusedParams p, Epi epi) {
  using C = Cfg<HS, O>;
  constexpr int NST = C::NST;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[NST];
  __shared__ __align__(8) uint64_t empty_bar[NST];
  __shared__ __align__(8) uint64_t p_full_bar[2];    // MMA1 of a P buffer retired (tcgen05.commit)
  __shared__ __align__(8) uint64_t p_ready_bar[2];   // the 4 rounding warps have rewritten the P buffer
  __shared__ __align__(8) uint64_t acc_full_bar;
  __shared__ uint32_t tmem_slot;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, b = blockIdx.y;
  const int kb1 = (p.N + BK - 1) / BK;
  const int NBLK = p.KS + 1;                         // weight segments per part: NB + 1 = KS + 2; NBLK = index of the input block
  if (threadIdx.x == 0) MCRN_TL(0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmS) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmXA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmXB) : "memory");
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
  if (threadIdx.x == 0) MCRN_TL(1);

  if (warp == 0) {
    if (lane == 0) {                                     // ===== TMA producer =====
      int it = 0;
      for_each_item<C::KB2>(p.KS, kb1, p.nparts, [&](int type, int k, int j, int part) {
        const int s = it % NST;
        if (it >= NST) mbar_wait_b(smem_u32(&empty_bar[s]), (((uint32_t)(it / NST)) & 1u) ^ 1u);
        if (it < 200) MCRN_TL(240 + it);
        const uint32_t fb = smem_u32(&full_bar[s]);
        const uint32_t a_dst = smem_base + (uint32_t)s * C::STAGE, b_dst = a_dst + C::A_SLOT;
        if (type == ITEM_P) {
          mbar_expect_tx(fb, C::A_SLOT + (uint32_t)HS * BK * 4);
          tma_load_4d(a_dst, &tmS, fb, j * BK, m0, k, 0);                       // S_k[m0.., 32 j..]  (K-major)
#pragma unroll
          for (int q = 0; q < HS / 32; ++q)                                      // X[32 j.., b, 32 q..]  (MN-major slabs)
            tma_load_4d(b_dst + q * SLAB_BYTES, &tmXB, fb, 32 * q, b, j * BK, 0);
        } else {
          const int wseg = (type == ITEM_SS ? k : 1 + k) + part * (NBLK + 1);
          if (type == ITEM_SS) {
            mbar_expect_tx(fb, C::A_SLOT + C::B_SLOT);
            tma_load_4d(a_dst, &tmXA, fb, j * BK, b, m0, k);                    // XP[blk k][m0.., b, 32 j..]  (K-major)
          } else {
            mbar_expect_tx(fb, C::B_SLOT);
          }
#pragma unroll
          for (int q = 0; q < O / 32; ++q)                                       // W[wseg][32 j.., 32 q..]  (MN-major slabs)
            tma_load_4d(b_dst + q * SLAB_BYTES, &tmW, fb, 32 * q, j * BK, wseg, 0);
        }
        ++it;
      });
      MCRN_TL(232);
    }
  } else if (warp == 1) {
    if (lane == 0) {                                     // ===== MMA issuer =====
      constexpr uint32_t idesc1 = make_idesc<true, false, HS>();
      constexpr uint32_t idesc2 = make_idesc<true, false, O>();
      int it = 0;
      bool acc_on = false;
      for_each_item<C::KB2>(p.KS, kb1, p.nparts, [&](int type, int k, int j, int part) {
        const int s = it % NST;
        mbar_wait_b(smem_u32(&full_bar[s]), ((uint32_t)(it / NST)) & 1u);
        tcgen05_fence_after();
        if (it < 200) MCRN_TL(2 + it);
        const uint32_t a_addr = smem_base + (uint32_t)s * C::STAGE, b_addr = a_addr + C::A_SLOT;
        const uint32_t pbuf = tmem_base + ((k & 1) ? C::TM_P1 : C::TM_P0);
        if (type == ITEM_P) {
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024, 2);
            const uint64_t bd = make_smem_desc(b_addr + kk * 1024, SLAB_BYTES, 512, 1);
            tcgen05_mma_tf32(pbuf, ad, bd, idesc1, (j > 0 || kk > 0) ? 1u : 0u);
          }
          tcgen05_commit(smem_u32(&empty_bar[s]));
          if (j == kb1 - 1) tcgen05_commit(smem_u32(&p_full_bar[k & 1]));
        } else if (type == ITEM_SS) {
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t ad = make_smem_desc(a_addr + kk * 32, 16, 1024, 2);
            const uint64_t bd = make_smem_desc(b_addr + kk * 1024, SLAB_BYTES, 512, 1);
            tcgen05_mma_tf32(tmem_base + C::TM_ACC, ad, bd, idesc2, (acc_on || kk > 0) ? 1u : 0u);
          }
          acc_on = true;
          tcgen05_commit(smem_u32(&empty_bar[s]));
        } else {
          if (j == 0 && part == 0) {                     // P_k has been rounded in place by the epilogue warps
            mbar_wait_b(smem_u32(&p_ready_bar[k & 1]), ((uint32_t)(k >> 1)) & 1u);
            tcgen05_fence_after();
          }
#pragma unroll
          for (int kk = 0; kk < BK / 8; ++kk) {
            const uint64_t bd = make_smem_desc(b_addr + kk * 1024, SLAB_BYTES, 512, 1);
            tcgen05_mma_tf32_ts(tmem_base + C::TM_ACC, pbuf + (uint32_t)(j * BK + kk * 8), bd, idesc2, (acc_on || kk > 0) ? 1u : 0u);
          }
          acc_on = true;
          tcgen05_commit(smem_u32(&empty_bar[s]));
        }
        ++it;
      });
      tcgen05_commit(smem_u32(&acc_full_bar));
      MCRN_TL(233);
    }
  } else {                                               // ===== rounding + epilogue warps =====
    const int quarter = warp & 3;                        // TMEM lane quarter this warp may access
    float* scr = reinterpret_cast<float*>(smem_al + (size_t)NST * C::STAGE) + (warp - 2) * (32 * 36);
    const int cq = (lane & 7) * 4, r0 = lane >> 3;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int node0 = m0 + quarter * 32;
    // ---- P_k: TF32 round-to-nearest in place; training: store the rounded block for the backward ----
    for (int k = 0; k < p.KS; ++k) {
      mbar_wait_b(smem_u32(&p_full_bar[k & 1]), ((uint32_t)(k >> 1)) & 1u);
      tcgen05_fence_after();
      if (warp == 2 && lane == 0 && k < 5) MCRN_TL(210 + 4 * k);
      const uint32_t pbuf = tmem_base + ((k & 1) ? C::TM_P1 : C::TM_P0) + lane_off;
#pragma unroll 1
      for (int c = 0; c < HS / 32; ++c) {
        float v[32];
        tmem_ld_32x32b_x32(pbuf + (uint32_t)(c * 32), v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = tf32_rn(v[i]);
        tmem_st_32x32b_x32(pbuf + (uint32_t)(c * 32), v);
        if (p.xp_save != nullptr && node0 < p.N) {
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(&scr[lane * 36 + i]) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
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
      if (warp == 2 && lane == 0 && k < 5) MCRN_TL(211 + 4 * k);
    }
    // ---- epilogue: accumulator -> gate / update math (rows = (node, b), all O columns) ----
    mbar_wait_b(smem_u32(&acc_full_bar), 0);
    tcgen05_fence_after();
    if (warp == 2 && lane == 0) MCRN_TL(230);
    if (node0 < p.N) {
#pragma unroll 1
      for (int c = 0; c < O / 32; ++c) {
        float v[32];
        tmem_ld_32x32b_x32(tmem_base + C::TM_ACC + lane_off + (uint32_t)(c * 32), v);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(&scr[lane * 36 + i]) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        __syncwarp();
        const int col = c * 32 + cq;
        bool done = false;
        if constexpr (Epi::NP > 0) {
          if (epi.fast4(0, 0, col)) {
            constexpr int RB = Epi::NP <= 2 ? 8 : (Epi::NP <= 4 ? 4 : 2);
#pragma unroll
            for (int b0 = 0; b0 < 8; b0 += RB) {
              float4 pre[RB][Epi::NP];
#pragma unroll
              for (int i = 0; i < RB; ++i) {
                const int node = node0 + r0 + 4 * (b0 + i);
                if (node < p.N) epi.load4(0, node * p.B + b, col, pre[i]);
              }
#pragma unroll
              for (int i = 0; i < RB; ++i) {
                const int rr = r0 + 4 * (b0 + i), node = node0 + rr;
                if (node < p.N) {
                  const float4 t = *reinterpret_cast<const float4*>(&scr[rr * 36 + cq]);
                  const float a4[4] = {t.x, t.y, t.z, t.w};
                  epi.fin4(0, node * p.B + b, col, pre[i], a4);
                }
              }
            }
            done = true;
          }
        }
        if (!done) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = r0 + 4 * i, node = node0 + rr;
            if (node < p.N) {
              const float4 t = *reinterpret_cast<const float4*>(&scr[rr * 36 + cq]);
              float a4[4] = {t.x, t.y, t.z, t.w};
              epi.template apply<4>(0, node * p.B + b, col, 4, a4);
            }
          }
        }
      }
    }
  }
  if (warp == 2 && lane == 0) MCRN_TL(231);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::TMEM_COLS) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------
// Which (HS, O) the fused kernel is instantiated for.
static inline bool fused_eligible(int N, int B, int HS, int O, const void* S, const void* xp, const void* w) {
  if (!(HS == 64 || HS == 128) || !(O == HS || O == 2 * HS)) return false;
  if ((reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(w)) & 15) return false;
  return N >= 1 && B >= 1 && B <= 65535;
}

// S: [KS][N][ldS] TF32-rounded supports.  xp: XP buffer [KS+2][R][HS] of this AGCN (block 0 = state, block KS+1 = input
// block; blocks 1..KS are written here when save != 0).  w: folded weights [nparts][KS+2][HS][O].
template <int HS, int O, class Epi>
int launch_agcn_fused(int N, int B, int KS, int ldS, const float* S, float* xp, const float* w, int nparts, int save,
                      const Epi& epi, cudaStream_t st) {
  using C = Cfg<HS, O>;
  const int64_t R = (int64_t)N * B;
  CUtensorMap tS, tXA, tXB, tW;
  {
    uint64_t dims[4] = {(uint64_t)N, (uint64_t)N, (uint64_t)KS, 1};
    uint64_t str[3] = {(uint64_t)ldS * 4, (uint64_t)N * ldS * 4, (uint64_t)KS * N * ldS * 4};
    uint32_t box[4] = {32, BM, 1, 1};
    MCRN_TRY(encode_tensor_map(&tS, S, dims, str, box, false));
  }
  {
    uint64_t dims[4] = {(uint64_t)HS, (uint64_t)B, (uint64_t)N, (uint64_t)(KS + 2)};
    uint64_t str[3] = {(uint64_t)HS * 4, (uint64_t)B * HS * 4, (uint64_t)R * HS * 4};
    uint32_t boxa[4] = {32, 1, BM, 1};
    MCRN_TRY(encode_tensor_map(&tXA, xp, dims, str, boxa, false));
    uint32_t boxb[4] = {32, 1, BK, 1};
    MCRN_TRY(encode_tensor_map(&tXB, xp, dims, str, boxb, true));
  }
  {
    uint64_t dims[4] = {(uint64_t)O, (uint64_t)HS, (uint64_t)(nparts * (KS + 2)), 1};
    uint64_t str[3] = {(uint64_t)O * 4, (uint64_t)HS * O * 4, (uint64_t)nparts * (KS + 2) * HS * O * 4};
    uint32_t box[4] = {32, BK, 1, 1};
    MCRN_TRY(encode_tensor_map(&tW, w, dims, str, box, true));
  }
  FusedParams p;
  p.N = N; p.B = B; p.KS = KS; p.nparts = nparts;
  p.xp_save = save ? xp : nullptr;
  p.blk_stride = R * HS;
  p.dbg = nullptr;
  if (g_dbg_timeline != nullptr) {
    if (g_dbg_which < 0 || g_dbg_count == g_dbg_which) p.dbg = g_dbg_timeline;
    ++g_dbg_count;
  }
  auto kern = agcn_fused_kernel<HS, O, Epi>;
  static bool attr_set = false;
  if (!attr_set) {
    MCRN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    attr_set = true;
  }
  dim3 grid(ceil_div(N, BM), B, 1);
  const int pi = prof_begin(prof_class(0, HS, O == HS ? 1 : 0), st);
  MCRN_LAUNCH(kern, grid, THREADS, C::SMEM, st, tS, tXA, tXB, tW, p, epi);
  prof_end(pi, st);
  return MCRN_OK;
}

}  // namespace fused
}  // namespace mcrn
