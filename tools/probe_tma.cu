// Per-SM TMA ingest probe (experiment, not part of the library): one CTA per SM streams fp16 boxes [64 halves x ROWS]
// (128-byte swizzle) from an L2-resident buffer through an NST-deep mbarrier ring; the consumer either releases the slot
// at once (MMA = 0) or issues ROWS/128 x 4 tcgen05.mma (M = 128, N = NMMA, K = 16) per box and releases it with
// tcgen05.commit.  Prints bytes / clock / SM as a function of the box size, the ring depth and the grid size.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_tma tools/probe_tma.cu -lcuda
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)(16 >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct P { int items, rows, nst, mma, nmma, total_rows; long long* out; };

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm, P p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[16], empty_bar[16];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t box_bytes = (uint32_t)p.rows * 128;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nst; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    int row = (blockIdx.x * 977) % p.total_rows;
    for (int it = 0; it < p.items; ++it) {
      const int s = it % p.nst;
      if (it >= p.nst) mbar_wait(smem_u32(&empty_bar[s]), (((uint32_t)(it / p.nst)) & 1u) ^ 1u);
      const uint32_t fb = smem_u32(&full_bar[s]);
      mbar_expect_tx(fb, box_bytes);
      for (int r = 0; r < p.rows; r += 128) {          // boxes of <= 256 rows: issue 128-row pieces
        if (row + 128 > p.total_rows) row = 0;
        tma_load_2d(base + (uint32_t)s * box_bytes + (uint32_t)r * 128, &tm, fb, 0, row);
        row += 128;
      }
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(p.nmma >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int it = 0; it < p.items; ++it) {
      const int s = it % p.nst;
      mbar_wait(smem_u32(&full_bar[s]), ((uint32_t)(it / p.nst)) & 1u);
      if (p.mma) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a = base + (uint32_t)s * box_bytes;
        for (int r = 0; r < p.mma; ++r)
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t ad = make_desc(a + kk * 32), bd = make_desc(a + kk * 32);
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                         ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
          }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
      } else {
        mbar_arrive(smem_u32(&empty_bar[s]));
      }
    }
    if (p.mma) {      // drain: wait for the last commit of every slot
      for (int s = 0; s < p.nst && s < p.items; ++s) {
        const int last = ((p.items - 1 - s) / p.nst) * p.nst + s;
        mbar_wait(smem_u32(&empty_bar[s]), ((uint32_t)(last / p.nst)) & 1u);
      }
    }
    p.out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
}

int main() {
  const int total_rows = 128 * 256;                       // 4 MB buffer: L2 resident
  __half* buf;
  CK(cudaMalloc(&buf, (size_t)total_rows * 128));
  CK(cudaMemset(buf, 0, (size_t)total_rows * 128));
  long long* out;
  CK(cudaMalloc(&out, 148 * sizeof(long long)));
  CUtensorMap tm;
  cuuint64_t gdim[2] = {64, (cuuint64_t)total_rows};
  cuuint64_t gstr[1] = {128};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t es[2] = {1, 1};
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024));
  printf("rows  KB/box nst  inflightKB grid mma(nmma) cycles/item  B/clk/SM (median CTA)\n");
  const int rows_l[] = {128, 256, 384};
  const int grids[] = {16, 128, 148};
  for (int mma = 0; mma < 3; ++mma)
    for (int gi = 0; gi < 3; ++gi)
      for (int ri = 0; ri < 3; ++ri)
        for (int nst = 2; nst <= 12; nst += (nst < 4 ? 1 : 2)) {
          const int rows = rows_l[ri];
          if ((size_t)nst * rows * 128 > 200 * 1024) continue;
          P p{400, rows, nst, mma == 0 ? 0 : rows / 128, mma == 2 ? 256 : 128, total_rows, out};
          if (mma == 2 && rows < 256) continue;           // N = 256 needs a 256-row B tile in the slot
          for (int rep = 0; rep < 2; ++rep) {
            probe<<<grids[gi], 128, 200 * 1024 + 1024>>>(tm, p);
            CK(cudaDeviceSynchronize());
          }
          long long h[148];
          CK(cudaMemcpy(h, out, grids[gi] * sizeof(long long), cudaMemcpyDeviceToHost));
          // median
          for (int i = 0; i < grids[gi]; ++i)
            for (int j = i + 1; j < grids[gi]; ++j)
              if (h[j] < h[i]) { long long t = h[i]; h[i] = h[j]; h[j] = t; }
          const double cyc = (double)h[grids[gi] / 2] / p.items;
          printf("%4d  %5d  %3d  %6d    %4d  %d(%d)   %8.1f   %6.1f\n", rows, rows * 128 / 1024, nst, nst * rows * 128 / 1024, grids[gi],
                 p.mma, p.nmma, cyc, rows * 128.0 / cyc);
        }
  return 0;
}
