// TMA request-rate probe 2 (experiment): ITEMS items are issued back to back into ITEMS distinct slots (no slot reuse, no
// consumer handshake); a second thread polls the full barriers in order and stamps when each item has landed.
// Variants: boxes per item, rows per box (128 / 256), descriptor in kernel params vs global memory, one or two issuing
// threads, tensor (tiled) loads vs plain 1-D bulk copies of the same bytes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_tma2 tools/probe_tma2.cu -lcuda
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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n.reg .pred P1;\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int MAXI = 12;
struct P { int items, boxes, rows, mode, nthr, total_rows, poll; const CUtensorMap* tm_g; const uint8_t* buf; long long* out; };
// mode 0: tensor load, descriptor = kernel param; 1: tensor load, descriptor in global memory; 2: 1-D bulk copies

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm128, const __grid_constant__ CUtensorMap tm256, P p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[MAXI];
  __shared__ long long t_issue[MAXI], t_land[MAXI];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t box_bytes = (uint32_t)p.rows * 128, item_bytes = box_bytes * p.boxes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.items; ++s) mbar_init(smem_u32(&full_bar[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (warp < p.nthr && lane == 0) {
    const CUtensorMap* tm = p.mode == 1 ? p.tm_g + (p.rows == 256 ? 1 : 0) : (p.rows == 256 ? &tm256 : &tm128);
    int row = ((blockIdx.x * 977) % (p.total_rows / 256)) * 256;
    for (int it = warp; it < p.items; it += p.nthr) {
      const uint32_t fb = smem_u32(&full_bar[it]);
      t_issue[it] = clock64() - t0;
      mbar_expect_tx(fb, item_bytes);
      for (int b = 0; b < p.boxes; ++b) {
        const int r = (row + (it * p.boxes + b) * p.rows) % (p.total_rows - 256);
        const uint32_t dst = base + (uint32_t)it * item_bytes + (uint32_t)b * box_bytes;
        if (p.mode == 2) bulk_load_1d(dst, p.buf + (size_t)r * 128, box_bytes, fb);
        else tma_load_2d(dst, tm, fb, 0, r);
      }
    }
  } else if (warp == 3 && lane == 0) {
    for (int it = 0; it < p.items; ++it) {
      if (p.poll) { while (!mbar_test(smem_u32(&full_bar[it]), 0)) {} }
      else mbar_wait(smem_u32(&full_bar[it]), 0);
      t_land[it] = clock64() - t0;
    }
  }
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x < p.items) { p.out[threadIdx.x] = t_issue[threadIdx.x]; p.out[MAXI + threadIdx.x] = t_land[threadIdx.x]; }
}

int main() {
  const int total_rows = 128 * 256;
  uint8_t* buf;
  CK(cudaMalloc(&buf, (size_t)total_rows * 128));
  CK(cudaMemset(buf, 0, (size_t)total_rows * 128));
  long long* out;
  CK(cudaMalloc(&out, 2 * MAXI * sizeof(long long)));
  CUtensorMap tm[2];
  for (int i = 0; i < 2; ++i) {
    cuuint64_t gdim[2] = {64, (cuuint64_t)total_rows};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)(i ? 256 : 128)};
    cuuint32_t es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  CUtensorMap* tm_g;
  CK(cudaMalloc(&tm_g, 2 * sizeof(CUtensorMap)));
  CK(cudaMemcpy(tm_g, tm, 2 * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024 + 1024));
  struct V { int items, boxes, rows, mode, nthr, poll; const char* name; };
  const V vs[] = {
      {12, 1, 128, 0, 1, 0, "tensor, param desc, 1 box x128 rows per item"},
      {6, 2, 128, 0, 1, 0, "tensor, param desc, 2 boxes x128"},
      {6, 1, 256, 0, 1, 0, "tensor, param desc, 1 box x256"},
      {3, 4, 128, 0, 1, 0, "tensor, param desc, 4 boxes x128"},
      {3, 2, 256, 0, 1, 0, "tensor, param desc, 2 boxes x256"},
      {12, 1, 128, 1, 1, 0, "tensor, GLOBAL desc, 1 box x128"},
      {12, 1, 128, 0, 2, 0, "tensor, param desc, 1 box x128, TWO issuing threads"},
      {12, 1, 128, 0, 1, 1, "tensor, param desc, 1 box x128, consumer polls test_wait"},
      {12, 1, 128, 2, 1, 0, "1-D bulk copy 16 KB per item"},
      {6, 1, 256, 2, 1, 0, "1-D bulk copy 32 KB per item"},
      {3, 2, 256, 2, 1, 0, "1-D bulk copy 2 x 32 KB per item"},
  };
  for (int grid : {1, 148}) {
    for (const V& v : vs) {
      P p{v.items, v.boxes, v.rows, v.mode, v.nthr, total_rows, v.poll, tm_g, buf, out};
      for (int rep = 0; rep < 3; ++rep) {
        probe<<<grid, 128, 200 * 1024 + 1024>>>(tm[0], tm[1], p);
        CK(cudaDeviceSynchronize());
      }
      long long h[2 * MAXI];
      CK(cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost));
      printf("grid %3d | %-58s | issue:", grid, v.name);
      for (int i = 0; i < v.items; ++i) printf(" %lld", h[i]);
      printf(" | landed:");
      for (int i = 0; i < v.items; ++i) printf(" %lld", h[MAXI + i]);
      const double per = v.items > 1 ? (double)(h[MAXI + v.items - 1] - h[MAXI]) / (v.items - 1) : 0.0;
      printf(" | %.0f cyc/item, %.1f B/clk\n", per, per > 0 ? v.boxes * v.rows * 128.0 / per : 0.0);
    }
  }
  return 0;
}
