// SIMT fp32 GEMM: exact-fp32 engine used (a) for shapes the tcgen05 path does not take
// (stride alignment below 16 B, tiny problems), (b) as the on-device cross-check of the
// tensor-core kernels in tests, and (c) when mcrn_set_engine(1) forces it.
// 64x64x32 tiles, 256 threads, 4x4 register micro-tiles, register-prefetched k-tiles, operands through generic strides.
#pragma once

#include "gemm.cuh"

namespace mcrn {

constexpr int SIMT_BM = 64, SIMT_BN = 64, SIMT_BK = 32, SIMT_THREADS = 256;

// 64x64x32 tiles; the global loads of k-tile i+1 are issued into registers before the FMAs of k-tile i, so the load
// latency of the (short, latency-bound) problems this engine serves overlaps the arithmetic.
template <class Epi>
__global__ void __launch_bounds__(SIMT_THREADS) gemm_simt_kernel(GemmDesc g, Epi epi) {
  __shared__ float As[SIMT_BK][SIMT_BM + 4];
  __shared__ float Bs[SIMT_BK][SIMT_BN + 4];
  const int tid = threadIdx.x;
  const int bz = blockIdx.z / g.splits, split = blockIdx.z - bz * g.splits;
  const int m0 = blockIdx.y * SIMT_BM, n0 = blockIdx.x * SIMT_BN;
  const float* A = g.A + (int64_t)bz * g.a_batch;
  const float* B = g.B + (int64_t)bz * g.b_batch;
  const int ktiles = (g.Kseg + SIMT_BK - 1) / SIMT_BK;
  const int total = g.total_segs() * ktiles;
  const int per = (total + g.splits - 1) / g.splits;
  const int it0 = split * per, it1 = min(total, it0 + per);
  const bool a_kc = (g.a_k == 1), b_nc = (g.b_n == 1);
  const int tx = tid & 15, ty = tid >> 4;
  constexpr int EPT = SIMT_BM * SIMT_BK / SIMT_THREADS;      // 8 elements of each operand per thread and k-tile
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[EPT], rb[EPT];
  auto fetch = [&](int it) {
    const int seg = it / ktiles, k0 = (it - seg * ktiles) * SIMT_BK;
    const float* As_g = A + (int64_t)g.seg_a(seg) * g.a_seg;
    const float* Bs_g = B + (int64_t)g.seg_b(seg) * g.b_seg;
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int idx = tid + e * SIMT_THREADS;     // 0..2047
      int mm, kk;
      if (a_kc) { mm = idx >> 5; kk = idx & 31; } else { kk = idx >> 6; mm = idx & 63; }
      const int gm = m0 + mm, gk = k0 + kk;
      ra[e] = (gm < g.M && gk < g.Kseg) ? __ldg(As_g + (int64_t)gm * g.a_row + (int64_t)gk * g.a_k) : 0.f;
      int nn, kb;
      if (b_nc) { kb = idx >> 6; nn = idx & 63; } else { nn = idx >> 5; kb = idx & 31; }
      const int gn = n0 + nn, gkb = k0 + kb;
      rb[e] = (gn < g.N && gkb < g.Kseg) ? __ldg(Bs_g + (int64_t)gkb * g.b_k + (int64_t)gn * g.b_n) : 0.f;
    }
  };
  if (it0 < it1) fetch(it0);
  for (int it = it0; it < it1; ++it) {
#pragma unroll
    for (int e = 0; e < EPT; ++e) {
      const int idx = tid + e * SIMT_THREADS;
      int mm, kk;
      if (a_kc) { mm = idx >> 5; kk = idx & 31; } else { kk = idx >> 6; mm = idx & 63; }
      As[kk][mm] = ra[e];
      int nn, kb;
      if (b_nc) { kb = idx >> 6; nn = idx & 63; } else { nn = idx >> 5; kb = idx & 31; }
      Bs[kb][nn] = rb[e];
    }
    __syncthreads();
    if (it + 1 < it1) fetch(it + 1);                // in flight while this tile is multiplied
#pragma unroll
    for (int kk = 0; kk < SIMT_BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (it1 <= it0 && g.splits > 1) return;   // empty split contributes nothing
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    int n = n0 + tx * 4;
    if (m < g.M && n < g.N) epi.template apply<4>(bz, m, n, min(4, g.N - n), acc[i]);
  }
}

template <class Epi>
int gemm_simt(const GemmDesc& g, const Epi& epi, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0) return MCRN_OK;
  dim3 grid(ceil_div(g.N, SIMT_BN), ceil_div(g.M, SIMT_BM), g.nbatch * g.splits);
  MCRN_LAUNCH(gemm_simt_kernel<Epi>, grid, SIMT_THREADS, 0, stream, g, epi);
  return MCRN_OK;
}

}  // namespace mcrn
