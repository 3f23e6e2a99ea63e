// The supports prologue (model/MegaCRN.py:169-173 + the hoisted Chebyshev set of :19-23) and its backward as ONE cooperative
// launch each, for the graph sizes where they are latency-bound (N <= 512, cheb_k = 3: METR-LA, PEMS-BAY).
//
// At N = 207 the per-stage version is 9 (forward) + 14 (backward) launches of ~5-17 us each -- tiny exact-fp32 GEMMs and
// row kernels whose only cost is launch + memory latency -- and all of them sit on the critical path at the head / tail of the
// training step (~200 us of a 3.2 ms step).  Here every phase is a grid-strided loop over matrix rows (one row per block
// iteration, exact fp32 FMA, operands through shared memory / L1) and the phases are separated by grid.sync().
//
//   forward :  E_i = We_i Mem  |  L_i = E_i E_j^T, g_i = softmax(relu(L_i))  |  T2_i = 2 g_i g_i - I
//   backward:  dg_i = dT1_i + 2 (dT2_i g_i^T + g_i^T dT2_i)  |  dL_i = relu'softmax'(dg_i)  |
//              dE1 = (dLa + dLb^T) E2, dE2 = (dLa^T + dLb) E1  |  dWe_i = dE_i Mem^T, dMem += We_i^T dE_i
// Outputs are exactly the buffers the per-stage path fills (S, Sr, S16, E, L / the 4 gradient tensors), so the rest of the
// library does not know which path ran.  tests: test_supports_stage_matches_spec, every gradient-parity test (We1, We2, Memory).
#pragma once

#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include "small_kernels.cuh"

namespace mcrn {
namespace scoop {

namespace cg = cooperative_groups;
constexpr int THREADS = 256;

struct FwdArgs {
  const float *we1, *we2, *mem;      // [N][M], [N][M], [M][d]
  float *E1, *E2;                    // [N][d]
  float *L1, *L2;                    // [N][ld] logits (kept for the backward)
  float *S, *Sr;                     // [4][N][ld]: g1, T2(g1), g2, T2(g2); exact and TF32-rounded
  __half* S16;                       // [4][N][ld16] or null
  int N, M, d, ld, ld16;
};

__device__ __forceinline__ float block_sum(float v, float* sh) { return block_sum_256(v, sh); }

__global__ void __launch_bounds__(THREADS) k_supports_fwd_coop(FwdArgs a) {
  extern __shared__ float smem[];                    // one row of up to max(N, d, M) floats + reduction scratch
  __shared__ float red[8];
  cg::grid_group grid = cg::this_grid();
  const int N = a.N, M = a.M, d = a.d, ld = a.ld;
  const int64_t mat = (int64_t)N * ld;
  // ---- phase 1: E_i = We_i Mem                                               model/MegaCRN.py:169-170
  for (int r = blockIdx.x; r < 2 * N; r += gridDim.x) {
    const int i = r / N, n = r - i * N;
    const float* we = (i ? a.we2 : a.we1) + (int64_t)n * M;
    for (int m = threadIdx.x; m < M; m += THREADS) smem[m] = we[m];
    __syncthreads();
    for (int j = threadIdx.x; j < d; j += THREADS) {
      float s = 0.f;
      for (int m = 0; m < M; ++m) s = fmaf(smem[m], a.mem[(int64_t)m * d + j], s);
      (i ? a.E2 : a.E1)[(int64_t)n * d + j] = s;
    }
    __syncthreads();
  }
  grid.sync();
  // ---- phase 2: logits, relu, row softmax                                    :171-172
  for (int r = blockIdx.x; r < 2 * N; r += gridDim.x) {
    const int i = r / N, n = r - i * N;
    const float* Ei = (i ? a.E2 : a.E1) + (int64_t)n * d;
    const float* Ej = i ? a.E1 : a.E2;
    float* L = (i ? a.L2 : a.L1) + (int64_t)n * ld;
    float* ev = smem;                                // E_i[n][:]
    float* lg = smem + d;                            // logits of the row
    for (int j = threadIdx.x; j < d; j += THREADS) ev[j] = Ei[j];
    __syncthreads();
    // one warp per column m, lanes along the contraction index (coalesced reads of E_j[m][:])
    for (int m = threadIdx.x >> 5; m < N; m += THREADS / 32) {
      const float* e = Ej + (int64_t)m * d;
      float s = 0.f;
      for (int j = threadIdx.x & 31; j < d; j += 32) s = fmaf(ev[j], e[j], s);
      s = warp_sum(s);
      if ((threadIdx.x & 31) == 0) { L[m] = s; lg[m] = fmaxf(s, 0.f); }
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int m = threadIdx.x; m < N; m += THREADS) mx = fmaxf(mx, lg[m]);
    mx = block_max_256(mx, red);
    float sum = 0.f;
    for (int m = threadIdx.x; m < N; m += THREADS) sum += expf(lg[m] - mx);
    sum = block_sum(sum, red);
    const float inv = 1.0f / sum;
    float* g = a.S + (int64_t)(2 * i) * mat + (int64_t)n * ld;
    float* gr = a.Sr + (int64_t)(2 * i) * mat + (int64_t)n * ld;
    __half* g16 = a.S16 ? a.S16 + ((int64_t)(2 * i) * N + n) * a.ld16 : nullptr;
    for (int m = threadIdx.x; m < N; m += THREADS) {
      const float v = expf(lg[m] - mx) * inv;
      g[m] = v;
      gr[m] = tf32_rn(v);
      if (g16) g16[m] = __float2half_rn(v);
    }
    for (int m = N + threadIdx.x; m < ld; m += THREADS) { g[m] = 0.f; gr[m] = 0.f; L[m] = 0.f; }
    if (g16) for (int m = N + threadIdx.x; m < a.ld16; m += THREADS) g16[m] = __float2half_rn(0.f);
    __syncthreads();
  }
  grid.sync();
  // ---- phase 3: T2 = 2 g g - I                                               :21-22 (hoisted)
  for (int r = blockIdx.x; r < 2 * N; r += gridDim.x) {
    const int i = r / N, n = r - i * N;
    const float* g = a.S + (int64_t)(2 * i) * mat;
    for (int m = threadIdx.x; m < N; m += THREADS) smem[m] = g[(int64_t)n * ld + m];
    __syncthreads();
    float* t2 = a.S + (int64_t)(2 * i + 1) * mat + (int64_t)n * ld;
    float* t2r = a.Sr + (int64_t)(2 * i + 1) * mat + (int64_t)n * ld;
    __half* t16 = a.S16 ? a.S16 + ((int64_t)(2 * i + 1) * N + n) * a.ld16 : nullptr;
    for (int c = threadIdx.x; c < N; c += THREADS) {
      float s = 0.f;
#pragma unroll 16
      for (int m = 0; m < N; ++m) s = fmaf(smem[m], g[(int64_t)m * ld + c], s);      // independent L2 loads: keep 16 in flight
      const float v = 2.0f * s - (c == n ? 1.0f : 0.0f);
      t2[c] = v;
      t2r[c] = tf32_rn(v);
      if (t16) t16[c] = __float2half_rn(v);
    }
    for (int c = N + threadIdx.x; c < ld; c += THREADS) { t2[c] = 0.f; t2r[c] = 0.f; }
    if (t16) for (int c = N + threadIdx.x; c < a.ld16; c += THREADS) t16[c] = __float2half_rn(0.f);
    __syncthreads();
  }
}

struct BwdArgs {
  const float *we1, *we2, *mem;      // parameters
  const float *E1, *E2, *L1, *L2;    // saved by the forward
  const float* S;                    // exact supports [4][N][ld]
  const float* dS;                   // [4][N][ld] accumulated support gradients: dT1(g1), dT2(g1), dT1(g2), dT2(g2)
  float *dLa, *dLb;                  // scratch [N][ld]
  float *dE1, *dE2;                  // scratch [N][d]
  float *g_we1, *g_we2, *g_mem;      // outputs: [N][M], [N][M], [M][d] (g_mem is ACCUMULATED into with atomics)
  int N, M, d, ld;
};

__global__ void __launch_bounds__(THREADS) k_supports_bwd_coop(BwdArgs a) {
  extern __shared__ float smem[];                    // 3 rows of N floats; in the last phase rows + an [M][d] partial of dMem
  __shared__ float red[8];
  cg::grid_group grid = cg::this_grid();
  const int N = a.N, M = a.M, d = a.d, ld = a.ld;
  const int64_t mat = (int64_t)N * ld;
  // ---- phase 1: dg = dT1 + 2 (dT2 g^T + g^T dT2), then the softmax / relu backward of the row ----
  for (int r = blockIdx.x; r < 2 * N; r += gridDim.x) {
    const int i = r / N, n = r - i * N;
    const float* g = a.S + (int64_t)(2 * i) * mat;
    const float* dT1 = a.dS + (int64_t)(2 * i) * mat;
    const float* dT2 = a.dS + (int64_t)(2 * i + 1) * mat;
    float* row = smem;                               // dT2[n][:]
    float* col = smem + N;                           // g[:][n]
    for (int m = threadIdx.x; m < N; m += THREADS) { row[m] = dT2[(int64_t)n * ld + m]; col[m] = g[(int64_t)m * ld + n]; }
    __syncthreads();
    const float* L = (i ? a.L2 : a.L1) + (int64_t)n * ld;
    float* s1v = smem + 2 * N;                       // (dT2 g^T)[n][:]: one warp per column c, lanes along m (coalesced rows of g)
    for (int c = threadIdx.x >> 5; c < N; c += THREADS / 32) {
      const float* gc = g + (int64_t)c * ld;
      float s1 = 0.f;
      for (int m = threadIdx.x & 31; m < N; m += 32) s1 = fmaf(row[m], gc[m], s1);
      s1 = warp_sum(s1);
      if ((threadIdx.x & 31) == 0) s1v[c] = s1;
    }
    __syncthreads();
    float dot = 0.f;                                 // sum_c g[n][c] dg[n][c]
    float dgv[4];                                    // N <= 4 * THREADS
    int q = 0;
    for (int c = threadIdx.x; c < N; c += THREADS, ++q) {
      float s2 = 0.f;
#pragma unroll 16
      for (int m = 0; m < N; ++m) s2 = fmaf(col[m], dT2[(int64_t)m * ld + c], s2);     // (g^T dT2)[n][c], coalesced over c
      const float v = dT1[(int64_t)n * ld + c] + 2.0f * (s1v[c] + s2);
      dgv[q] = v;
      dot = fmaf(g[(int64_t)n * ld + c], v, dot);
    }
    dot = block_sum(dot, red);
    float* dL = (i ? a.dLb : a.dLa) + (int64_t)n * ld;
    q = 0;
    for (int c = threadIdx.x; c < N; c += THREADS, ++q)
      dL[c] = (L[c] > 0.f) ? g[(int64_t)n * ld + c] * (dgv[q] - dot) : 0.f;
    __syncthreads();
  }
  grid.sync();
  // ---- phase 2: dE1[n] = sum_m (dLa[n][m] + dLb[m][n]) E2[m],  dE2[n] = sum_m (dLa[m][n] + dLb[n][m]) E1[m] ----
  for (int r = blockIdx.x; r < 2 * N; r += gridDim.x) {
    const int i = r / N, n = r - i * N;
    float* v = smem;
    for (int m = threadIdx.x; m < N; m += THREADS)
      v[m] = i == 0 ? a.dLa[(int64_t)n * ld + m] + a.dLb[(int64_t)m * ld + n] : a.dLa[(int64_t)m * ld + n] + a.dLb[(int64_t)n * ld + m];
    __syncthreads();
    const float* E = i == 0 ? a.E2 : a.E1;
    for (int j = threadIdx.x; j < d; j += THREADS) {
      float s = 0.f;
#pragma unroll 16
      for (int m = 0; m < N; ++m) s = fmaf(v[m], E[(int64_t)m * d + j], s);
      (i == 0 ? a.dE1 : a.dE2)[(int64_t)n * d + j] = s;
    }
    __syncthreads();
  }
  grid.sync();
  // ---- phase 3: dWe_i[n][m] = dE_i[n] . Mem[m];  dMem[m][j] += sum_n We_i[n][m] dE_i[n][j] (block partial, then atomics) ----
  float* part = smem + max(3 * N, d + M);            // [M][d]
  for (int e = threadIdx.x; e < M * d; e += THREADS) part[e] = 0.f;
  __syncthreads();
  for (int r = blockIdx.x; r < 2 * N; r += gridDim.x) {
    const int i = r / N, n = r - i * N;
    const float* dE = (i ? a.dE2 : a.dE1) + (int64_t)n * d;
    const float* we = (i ? a.we2 : a.we1) + (int64_t)n * M;
    float* ev = smem;                                // dE_i[n][:]
    float* wv = smem + d;                            // We_i[n][:]
    for (int j = threadIdx.x; j < d; j += THREADS) ev[j] = dE[j];
    for (int m = threadIdx.x; m < M; m += THREADS) wv[m] = we[m];
    __syncthreads();
    for (int m = threadIdx.x; m < M; m += THREADS) {
      float s = 0.f;
      for (int j = 0; j < d; ++j) s = fmaf(ev[j], a.mem[(int64_t)m * d + j], s);
      (i ? a.g_we2 : a.g_we1)[(int64_t)n * M + m] = s;
    }
    for (int e = threadIdx.x; e < M * d; e += THREADS) {
      const int m = e / d, j = e - m * d;
      part[e] = fmaf(wv[m], ev[j], part[e]);         // each thread owns its entries of `part`: no race
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < M * d; e += THREADS) atomicAdd(a.g_mem + e, part[e]);
}

// Shapes the cooperative path takes: cheb_k = 3 (4 supports), N small enough that a row fits one block iteration and the
// O(N^3 / gridDim) serial work per block stays below the launch latencies it replaces.
static inline bool eligible(int N, int cheb_k, int d, int M) { return cheb_k == 3 && N <= 512 && d <= 1024 && M <= 1024 && M * d <= 8192; }

static inline int coop_grid(const void* kern, size_t smem) {
  int dev = 0, sms = 0, per = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, THREADS, smem) != cudaSuccess || per < 1) return 0;
  return sms * (per > 6 ? 6 : per);      // as many resident blocks as there are rows to work on: one row per block and phase
}

static inline int launch_fwd(const FwdArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)(a.N + a.d + a.M + 8) * sizeof(float);      // phase 2: E row + logits row
  int grid = coop_grid((const void*)k_supports_fwd_coop, smem);
  if (grid < 1) { set_error("cooperative supports kernel: no resident block"); return MCRN_ERR_CUDA; }
  if (grid > 2 * a.N) grid = 2 * a.N;
  FwdArgs arg = a;
  void* params[] = {&arg};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_supports_fwd_coop, dim3(grid), dim3(THREADS), params, smem, st);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) { set_error("cudaLaunchCooperativeKernel(k_supports_fwd_coop): %s", cudaGetErrorString(e)); return MCRN_ERR_CUDA; }
  return MCRN_OK;
}

static inline int launch_bwd(const BwdArgs& a, cudaStream_t st) {
  const size_t smem = (size_t)((3 * a.N > a.d + a.M ? 3 * a.N : a.d + a.M) + a.M * a.d + 8) * sizeof(float);
  int grid = coop_grid((const void*)k_supports_bwd_coop, smem);
  if (grid < 1) { set_error("cooperative supports-backward kernel: no resident block"); return MCRN_ERR_CUDA; }
  if (grid > 2 * a.N) grid = 2 * a.N;
  BwdArgs arg = a;
  void* params[] = {&arg};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_supports_bwd_coop, dim3(grid), dim3(THREADS), params, smem, st);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (e != cudaSuccess) { set_error("cudaLaunchCooperativeKernel(k_supports_bwd_coop): %s", cudaGetErrorString(e)); return MCRN_ERR_CUDA; }
  return MCRN_OK;
}

}  // namespace scoop
}  // namespace mcrn
