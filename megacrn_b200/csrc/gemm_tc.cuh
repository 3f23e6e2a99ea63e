// tcgen05 TF32 GEMM for sm_100a: TMA (cp.async.bulk.tensor) stages fp32 operand tiles into
// 128B-swizzled shared memory, one elected thread issues tcgen05.mma.kind::tf32 with the fp32
// accumulator in TMEM, four epilogue warps read it back with tcgen05.ld and run the fused
// epilogue functor (gemm.cuh).  One 128 x BN output tile per CTA, warp-specialised:
//   warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
//
// Operand storage (all four combinations are instantiated):
//   K-major  : K contiguous  -> tile rows of 32 fp32 = one 128-byte swizzle row
//   MN-major : M/N contiguous -> slabs of [32 k][32 m] (128-byte rows along M/N)
// Inputs must already hold TF32-representable values if round-to-nearest behaviour is wanted
// (the tensor core truncates the low 13 mantissa bits); producers round with cvt.rna (gemm.cuh).
#pragma once

#include <cuda.h>

#include "gemm.cuh"

namespace mcrn {
namespace tc {

constexpr int BM = 128, BK = 32, THREADS = 192;
constexpr uint32_t SLAB_BYTES = 32 * 128;       // [32 k-rows][128 B] slab of an MN-major operand

struct TcParams {
  int M, N, Kseg, nseg, nbatch, splits, a_batched, b_batched;
  uint8_t a_map[16], b_map[16];   // K-segment -> operand segment (hi/lo splits, block buffers)
  int b_sub_seg;                  // BSUB > 1: B sub-tile j reads segment b_map[seg] + j * b_sub_seg
  float* dbg;            // debug dump (mcrn_debug_tc_gemm): [0, STAGE floats) = raw smem stage 0 after TMA
};
extern float* g_dbg;     // host-side: non-null only inside mcrn_debug_tc_gemm
extern int g_pdl;        // MCRN_PDL=1: launch the GEMM kernels with programmatic dependent launch

// ---- PTX wrappers -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
//   K-major fp32 : layout SWIZZLE_128B (2): rows of 128 B, 16-byte chunks XOR (row % 8); SBO = 1024 B per 8 rows.
//   MN-major fp32: layout SWIZZLE_128B_BASE32B (1) -- the only MN-major layout the hardware accepts for 32-bit
//                  operands: rows (k) of 128 B = 32 m, 32-byte chunks XOR (k % 4); atom = 4 k-rows (512 B);
//                  LBO = stride between 32-wide M/N groups, SBO = stride between 4-row k groups.
//                  The matching TMA mode is CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=TF32 [7,10)=2, B=TF32 [10,13)=2,
// a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), N>>3 [17,23), M>>4 [24,29).
template <bool A_K, bool B_K, int BN>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((A_K ? 0u : 1u) << 15) | ((B_K ? 0u : 1u) << 16) |
         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// Sub-tiling (CTA-level register blocking of the shared-memory traffic, which is what bounds these GEMMs:
// fp32 operands give a 128x128 tile only 32 FLOP per byte staged):
//   MSUB = 2 : two 128-row A sub-tiles share every B stage (CTA tile 256 x BN, two accumulators)
//   BSUB = 2 : two B sub-tiles (the TF32 hi and lo parts of the weights, `b_sub_seg` segments apart) multiply the
//              same A stage and accumulate into the same accumulator, so A is staged once for hi+lo.
template <int BN, int STAGES, int MSUB, int BSUB>
constexpr size_t smem_bytes() { return (size_t)STAGES * (MSUB * BM * BK * 4 + BSUB * BN * BK * 4) + 1024; }

template <bool A_K, bool B_K, int BN, int STAGES, int MSUB, int BSUB, class Epi>
__global__ void __launch_bounds__(THREADS, (smem_bytes<BN, STAGES, MSUB, BSUB>() <= 113 * 1024) ? 2 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p, Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[STAGES];
  __shared__ __align__(8) uint64_t empty_bar[STAGES];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_slot;
  constexpr uint32_t A_BYTES = BM * BK * 4, B_BYTES = BN * BK * 4, STAGE_BYTES = MSUB * A_BYTES + BSUB * B_BYTES;
  constexpr uint32_t TMEM_COLS = MSUB * BN;
  static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns: power of two");
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bz = blockIdx.z / p.splits, split = blockIdx.z - bz * p.splits;
  const int m0 = blockIdx.y * (BM * MSUB), n0 = blockIdx.x * BN;
  const int kt = (p.Kseg + BK - 1) / BK;
  const int total = p.nseg * kt;
  const int per = (total + p.splits - 1) / p.splits;
  const int it0 = split * per;
  const int nit = max(0, min(total, it0 + per) - it0);

  // One pipeline stage of TMA loads for local k-iteration i (executed by the producer thread only).
  auto produce = [&](int i) {
    const int s = i % STAGES;
    const uint32_t fb = smem_u32(&full_bar[s]);
    mbar_expect_tx(fb, STAGE_BYTES);
    const int bza = bz * p.a_batched, bzb = bz * p.b_batched;
    const int it = it0 + i, seg = it / kt, k0 = (it - seg * kt) * BK;
    const int sa = p.a_map[seg], sb = p.b_map[seg];
    const uint32_t a_dst = smem_base + (uint32_t)s * STAGE_BYTES, b_dst = a_dst + MSUB * A_BYTES;
#pragma unroll
    for (int ms = 0; ms < MSUB; ++ms) {
      if (A_K) {
        tma_load_4d(a_dst + ms * A_BYTES, &tmA, fb, k0, m0 + ms * BM, sa, bza);
      } else {
#pragma unroll
        for (int j = 0; j < BM / 32; ++j)
          tma_load_4d(a_dst + ms * A_BYTES + j * SLAB_BYTES, &tmA, fb, m0 + ms * BM + 32 * j, k0, sa, bza);
      }
    }
#pragma unroll
    for (int bs = 0; bs < BSUB; ++bs) {
      const int sbb = sb + bs * p.b_sub_seg;
      if (B_K) {
        tma_load_4d(b_dst + bs * B_BYTES, &tmB, fb, k0, n0, sbb, bzb);
      } else {
#pragma unroll
        for (int j = 0; j < BN / 32; ++j)
          tma_load_4d(b_dst + bs * B_BYTES + j * SLAB_BYTES, &tmB, fb, n0 + 32 * j, k0, sbb, bzb);
      }
    }
  };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
#pragma unroll
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // Programmatic dependent launch: everything above overlapped the previous kernel's tail; its results are
    // needed from here on (no-op when the kernel was launched without the PDL attribute).
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // The first STAGES loads need no free-slot handshake: issue them before the TMEM allocation / CTA barrier.
    const int pre = nit < STAGES ? nit : STAGES;
    for (int i = 0; i < pre; ++i) produce(i);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // every CTA of this grid is resident or done once all have reached this point: let the next kernel's CTAs start
  // their own prologue (they block in griddepcontrol.wait until this grid has completed and flushed)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (nit > 0) {
    if (warp == 0) {
      if (lane == 0) {                                   // ===== TMA producer (stages 0..STAGES-1 already in flight) =====
        for (int i = STAGES; i < nit; ++i) {
          const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
          mbar_wait(smem_u32(&empty_bar[i % STAGES]), ph ^ 1u);
          produce(i);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {                                   // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc<A_K, B_K, BN>();
        for (int i = 0; i < nit; ++i) {
          const int s = i % STAGES;
          const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
          mbar_wait(smem_u32(&full_bar[s]), ph);
          tcgen05_fence_after();
          if (p.dbg != nullptr && i == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
            const float* src = reinterpret_cast<const float*>(smem_raw + (smem_base - smem_u32(smem_raw)));
            for (uint32_t q = 0; q < STAGE_BYTES / 4; ++q) p.dbg[q] = src[q];
            p.dbg[STAGE_BYTES / 4] = __uint_as_float(tmem_base);
            p.dbg[STAGE_BYTES / 4 + 1] = __uint_as_float(smem_base);
          }
          const uint32_t a_addr = smem_base + (uint32_t)s * STAGE_BYTES, b_addr = a_addr + MSUB * A_BYTES;
#pragma unroll
          for (int ms = 0; ms < MSUB; ++ms) {
#pragma unroll
            for (int bs = 0; bs < BSUB; ++bs) {
#pragma unroll
              for (int kk = 0; kk < BK / 8; ++kk) {      // UMMA_K = 8 for tf32
                const uint32_t aa = a_addr + ms * A_BYTES, bb = b_addr + bs * B_BYTES;
                const uint64_t ad = A_K ? make_smem_desc(aa + kk * 32, 16, 1024, 2) : make_smem_desc(aa + kk * 1024, SLAB_BYTES, 512, 1);
                const uint64_t bd = B_K ? make_smem_desc(bb + kk * 32, 16, 1024, 2) : make_smem_desc(bb + kk * 1024, SLAB_BYTES, 512, 1);
                tcgen05_mma_tf32(tmem_base + (uint32_t)(ms * BN), ad, bd, idesc, (i > 0 || kk > 0 || bs > 0) ? 1u : 0u);
              }
            }
          }
          tcgen05_commit(smem_u32(&empty_bar[s]));       // frees the smem slot when these MMAs retire
        }
        tcgen05_commit(smem_u32(&tmem_full_bar));        // accumulator(s) complete
      }
    } else {                                             // ===== epilogue warps =====
      const int quarter = warp & 3;                      // TMEM lane quarter this warp may read
      mbar_wait(smem_u32(&tmem_full_bar), 0);
      tcgen05_fence_after();
      // Each thread owns one accumulator ROW (TMEM lane).  Storing rows per thread would scatter every warp
      // store over 32 cache lines, so each 32x32 chunk is transposed through shared memory (the pipeline
      // stages are idle once the accumulator is complete) and the epilogue functor runs with the lanes on
      // consecutive columns of a row (8 lanes x 4 columns = one 128-byte line per row, 4 rows per instruction):
      // all its global loads/stores are 128-bit and fully coalesced.
      float* scr = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw))) + quarter * (32 * 36);
      const int cq = (lane & 7) * 4, r0 = lane >> 3;       // this lane: 4 consecutive columns of rows r0, r0+4, ...
#pragma unroll 1
      for (int ms = 0; ms < MSUB; ++ms) {
        const int mrow0 = m0 + ms * BM + quarter * 32;
        if (mrow0 >= p.M) break;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          float v[32];
          tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ms * BN + c * 32), v);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(&scr[lane * 36 + j]) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          __syncwarp();
          const int col = n0 + c * 32 + cq;
          if (col < p.N) {
            const int nv = min(4, p.N - col);
            const int rbase = mrow0 + r0;
            bool done = false;
            if constexpr (Epi::NP > 0) {
              // two-phase fast path: issue the global loads of a whole batch of rows, then compute and store
              if (nv == 4 && epi.fast4(bz, rbase, col)) {
                constexpr int RB = Epi::NP <= 2 ? 8 : (Epi::NP <= 4 ? 4 : 2);
#pragma unroll
                for (int b0 = 0; b0 < 8; b0 += RB) {
                  float4 pre[RB][Epi::NP];
#pragma unroll
                  for (int i = 0; i < RB; ++i) {
                    const int row = rbase + 4 * (b0 + i);
                    if (row < p.M) epi.load4(bz, row, col, pre[i]);
                  }
#pragma unroll
                  for (int i = 0; i < RB; ++i) {
                    const int rr = r0 + 4 * (b0 + i);
                    const int row = mrow0 + rr;
                    if (row < p.M) {
                      const float4 t = *reinterpret_cast<const float4*>(&scr[rr * 36 + cq]);
                      const float a4[4] = {t.x, t.y, t.z, t.w};
                      epi.fin4(bz, row, col, pre[i], a4);
                    }
                  }
                }
                done = true;
              }
            }
            if (!done) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rr = r0 + 4 * i;
                const int row = mrow0 + rr;
                if (row < p.M) {
                  const float4 t = *reinterpret_cast<const float4*>(&scr[rr * 36 + cq]);
                  float a4[4] = {t.x, t.y, t.z, t.w};
                  epi.template apply<4>(bz, row, col, nv, a4);
                }
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------
int encode_tensor_map(CUtensorMap* out, const float* base, const uint64_t dims[4], const uint64_t strides_bytes[3],
                      const uint32_t box[4], bool mn_major);
bool eligible(const GemmDesc& g);

template <bool A_K, bool B_K, int BN, int STAGES, int MSUB, int BSUB, class Epi>
int launch(const GemmDesc& g, const Epi& epi, cudaStream_t st) {
  CUtensorMap ta, tb;
  {
    uint64_t dims[4], str[3];
    uint32_t box[4] = {32, 1, 1, 1};
    if (A_K) { dims[0] = g.Kseg; dims[1] = g.M; str[0] = g.a_row * 4; box[1] = BM; }
    else { dims[0] = g.M; dims[1] = g.Kseg; str[0] = g.a_k * 4; box[1] = BK; }
    const int aseg = g.a_seg ? g.nseg_a() : 1;   // distinct A segments (with use_map: a_nseg must be set)
    dims[2] = aseg; str[1] = aseg > 1 ? g.a_seg * 4 : str[0] * dims[1];
    const int ab = g.a_batch ? g.nbatch : 1;
    dims[3] = ab; str[2] = ab > 1 ? g.a_batch * 4 : str[1] * dims[2];
    MCRN_TRY(encode_tensor_map(&ta, g.A, dims, str, box, !A_K));
  }
  {
    uint64_t dims[4], str[3];
    uint32_t box[4] = {32, 1, 1, 1};
    if (B_K) { dims[0] = g.Kseg; dims[1] = g.N; str[0] = g.b_n * 4; box[1] = BN; }
    else { dims[0] = g.N; dims[1] = g.Kseg; str[0] = g.b_k * 4; box[1] = BK; }
    const int bseg = g.b_seg ? g.nseg_b() * (BSUB > 1 ? BSUB : 1) : 1;
    dims[2] = bseg; str[1] = bseg > 1 ? g.b_seg * 4 : str[0] * dims[1];
    const int bb = g.b_batch ? g.nbatch : 1;
    dims[3] = bb; str[2] = bb > 1 ? g.b_batch * 4 : str[1] * dims[2];
    MCRN_TRY(encode_tensor_map(&tb, g.B, dims, str, box, !B_K));
  }
  TcParams p;
  p.M = g.M; p.N = g.N; p.Kseg = g.Kseg; p.nseg = g.nseg;
  for (int i = 0; i < 16; ++i) {
    p.a_map[i] = (uint8_t)((i < g.nseg && g.a_seg) ? g.seg_a(i) : 0);
    p.b_map[i] = (uint8_t)((i < g.nseg && g.b_seg) ? g.seg_b(i) : 0);
  }
  p.nbatch = g.nbatch; p.splits = g.splits;
  p.a_batched = g.a_batch ? 1 : 0;
  p.b_batched = g.b_batch ? 1 : 0;
  p.b_sub_seg = g.b_sub_seg;
  p.dbg = g_dbg;
  auto kern = gemm_tc_kernel<A_K, B_K, BN, STAGES, MSUB, BSUB, Epi>;
  constexpr size_t smem = smem_bytes<BN, STAGES, MSUB, BSUB>();
  static bool attr_set = false;
  if (!attr_set) {
    MCRN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM * MSUB), g.nbatch * g.splits);
  if (!g_pdl) {
    MCRN_LAUNCH(kern, grid, THREADS, smem, st, ta, tb, p, epi);
    return MCRN_OK;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, ta, tb, p, epi);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (le != cudaSuccess) { set_error("cudaLaunchKernelEx(gemm_tc_kernel) failed: %s", cudaGetErrorString(le)); return MCRN_ERR_CUDA; }
  return MCRN_OK;
}

template <class Epi>
int gemm_tc(const GemmDesc& g, const Epi& epi, cudaStream_t st) {
  const bool a_k = (g.a_k == 1), b_k = (g.b_k == 1);
  const bool wide = g.N > 64;
  if (a_k && b_k) return wide ? launch<true, true, 128, 3, 1, 1, Epi>(g, epi, st) : launch<true, true, 64, 4, 1, 1, Epi>(g, epi, st);
  if (a_k && !b_k) return wide ? launch<true, false, 128, 3, 1, 1, Epi>(g, epi, st) : launch<true, false, 64, 4, 1, 1, Epi>(g, epi, st);
  if (!a_k && b_k) return wide ? launch<false, true, 128, 3, 1, 1, Epi>(g, epi, st) : launch<false, true, 64, 4, 1, 1, Epi>(g, epi, st);
  return wide ? launch<false, false, 128, 3, 1, 1, Epi>(g, epi, st) : launch<false, false, 64, 4, 1, 1, Epi>(g, epi, st);
}

// Weight contraction of the AGCN forward (A K-major = XP blocks, B MN-major = [hi | lo] weights, b_sub = 2):
// 256 x BN CTA tiles, the hi and lo weight tiles share each A stage.
extern int g_hilo_cfg;    // tuning knob (MCRN_HILO_CFG): which sub-tiling the hi/lo contraction uses
template <class Epi>
int gemm_tc_hilo(const GemmDesc& g, const Epi& epi, cudaStream_t st) {
  if (g.N <= 64) return launch<true, false, 64, 4, 2, 2, Epi>(g, epi, st);
  switch (g_hilo_cfg) {
    case 1: return launch<true, false, 128, 2, 1, 2, Epi>(g, epi, st);   // 128 x 128, A shared by hi/lo, 2 CTAs/SM
    case 2: return launch<true, false, 128, 4, 1, 2, Epi>(g, epi, st);   // same, 4 stages, 1 CTA/SM
    default: return launch<true, false, 128, 3, 2, 2, Epi>(g, epi, st);  // 256 x 128, 1 CTA/SM
  }
}

}  // namespace tc
}  // namespace mcrn
