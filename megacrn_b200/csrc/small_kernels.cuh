// Non-GEMM kernels of the MegaCRN hot path: layout changes, row softmax, memory-bank
// attention (warp-shuffle reductions), projection, and the elementwise / reduction
// pieces of the BPTT backward.  All HBM-bound; coalesced on the channel axis.
#pragma once

#include <cuda_fp16.h>

#include "gemm.cuh"

namespace mcrn {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reductions for 256-thread blocks (result valid in every thread).
__device__ __forceinline__ float block_sum_256(float v, float* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = ((threadIdx.x & 31) < 8) ? sh[threadIdx.x & 31] : 0.f;
  t = warp_sum(t);
  return __shfl_sync(0xffffffffu, t, 0);
}
__device__ __forceinline__ float block_max_256(float v, float* sh) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = ((threadIdx.x & 31) < 8) ? sh[threadIdx.x & 31] : -INFINITY;
  t = warp_max(t);
  return __shfl_sync(0xffffffffu, t, 0);
}

// ---- supports prologue ------------------------------------------------------------
// g[row,:] = softmax(relu(L[row,:]))                     model/MegaCRN.py:171-172
__global__ void __launch_bounds__(256) k_relu_softmax_rows(const float* __restrict__ L, float* __restrict__ G,
                                                           float* __restrict__ Gr, int n, int ld) {
  __shared__ float sh[8];
  const float* l = L + (int64_t)blockIdx.x * ld;
  float* g = G + (int64_t)blockIdx.x * ld;
  float* gr = Gr + (int64_t)blockIdx.x * ld;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < n; j += 256) mx = fmaxf(mx, fmaxf(l[j], 0.f));
  mx = block_max_256(mx, sh);
  float s = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) s += expf(fmaxf(l[j], 0.f) - mx);
  s = block_sum_256(s, sh);
  float inv = 1.0f / s;
  for (int j = threadIdx.x; j < n; j += 256) {
    float v = expf(fmaxf(l[j], 0.f) - mx) * inv;
    g[j] = v;
    gr[j] = tf32_rn(v);
  }
  for (int j = n + threadIdx.x; j < ld; j += 256) { g[j] = 0.f; gr[j] = 0.f; }
}

// dL[row,:] = g*(dg - sum(g*dg)) * (L > 0)      softmax + relu backward, one row per block
__global__ void __launch_bounds__(256) k_relu_softmax_rows_bwd(const float* __restrict__ L, const float* __restrict__ G,
                                                               const float* __restrict__ dG, float* __restrict__ dL,
                                                               int n, int ld) {
  __shared__ float sh[8];
  int64_t o = (int64_t)blockIdx.x * ld;
  float s = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) s += G[o + j] * dG[o + j];
  s = block_sum_256(s, sh);
  for (int j = threadIdx.x; j < n; j += 256)
    dL[o + j] = (L[o + j] > 0.f) ? G[o + j] * (dG[o + j] - s) : 0.f;
}

// out[i][j] = a[i][j] + b[j][i]   (square n x n, leading dim ld)
__global__ void k_add_transpose(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                                int n, int ld) {
  __shared__ float tile[32][33];
  int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = bx + i, c = by + threadIdx.x;            // read b[r][c] with r in the x-block
    tile[i][threadIdx.x] = (r < n && c < n) ? b[(int64_t)r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = by + i, c = bx + threadIdx.x;
    if (r < n && c < n) out[(int64_t)r * ld + c] = a[(int64_t)r * ld + c] + tile[threadIdx.x][i];
  }
}

// ---- parameter re-packing ---------------------------------------------------------------------
// One AGCN's weights w [2*ck*(cin+hs), O] + bias [O]  ->  wall [S][NB+1][hs][O]  (tests/kernel_spec.py:
// fold_agcn_weights), S = 2 when split (TF32 hi part, then the TF32-rounded residual lo; hi + lo carries ~21
// mantissa bits, removing the static -- hence coherent over time and batch -- weight-rounding error), else 1.
//   blocks 0..NB-1 : state-channel rows; block 0 = the two identity blocks summed (model/MegaCRN.py:20)
//   block NB       : the "input block": row k*cin+ci = input-channel row ci of block k, row NB*cin = bias, rest 0
__global__ void k_fold_weights(const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ wall,
                               int cin, int hs, int O, int ck, int split) {
  const int NB = 1 + 2 * (ck - 1), c = cin + hs;
  const int64_t total = (int64_t)(NB + 1) * hs * O;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % O);
    const int row = (int)((i / O) % hs);
    const int blk = (int)(i / ((int64_t)O * hs));
    int src_blk = blk, cc = cin + row;             // state row of block blk
    float v = 0.f;
    bool have = true;
    if (blk == NB) {
      if (row < NB * cin) { src_blk = row / cin; cc = row % cin; }
      else { have = false; if (row == NB * cin) v = bias[o]; }
    }
    if (have) {
      if (src_blk == 0) {
        v = w[((int64_t)0 * c + cc) * O + o] + w[((int64_t)ck * c + cc) * O + o];
      } else {
        int g = (src_blk - 1) / (ck - 1), k = 1 + (src_blk - 1) % (ck - 1);
        v = w[((int64_t)(g * ck + k) * c + cc) * O + o];
      }
    }
    if (!split) {
      wall[i] = v;
    } else {
      float hi = tf32_rn(v);
      wall[i] = hi;
      wall[total + i] = tf32_rn(v - hi);
    }
  }
}

// inverse for gradients: dwall [NB+1][hs][O] -> dw [2*ck*c, O] (both identity blocks get block 0) and dbias [O]
__global__ void k_unfold_grads(const float* __restrict__ dwall, float* __restrict__ dw, float* __restrict__ dbias,
                               int cin, int hs, int O, int ck) {
  const int NB = 1 + 2 * (ck - 1), c = cin + hs;
  const int64_t total = (int64_t)2 * ck * c * O;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total + O; i += (int64_t)gridDim.x * blockDim.x) {
    if (i >= total) {
      int o = (int)(i - total);
      dbias[o] = dwall[((int64_t)NB * hs + NB * cin) * O + o];
      continue;
    }
    int o = (int)(i % O);
    int cc = (int)((i / O) % c);
    int kk = (int)(i / ((int64_t)O * c));      // 0..2ck-1
    int g = kk / ck, k = kk % ck;
    int blk = (k == 0) ? 0 : 1 + g * (ck - 1) + (k - 1);
    dw[i] = (cc < cin) ? dwall[((int64_t)NB * hs + blk * cin + cc) * O + o]
                       : dwall[((int64_t)blk * hs + (cc - cin)) * O + o];
  }
}

// Input block of one step (block NB of both XP buffers of the step): column k*cin+ci = XPin[k][node][b][ci]
// (TF32-rounded when rnd), column NB*cin = 1 (multiplies the bias row), remaining columns 0.
__global__ void k_build_input_block(const float* __restrict__ xpin, int64_t xp_k, int64_t xp_n, int NB, int cin,
                                    int B, int64_t R, int hs, int rnd, float* __restrict__ ib_g,
                                    float* __restrict__ ib_u, __half* __restrict__ ib16) {
  const int64_t total = R * hs;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / hs;
    const int j = (int)(i - row * hs);
    float v = 0.f;
    if (j < NB * cin) {
      const int k = j / cin, ci = j - k * cin;
      const int node = (int)(row / B), b = (int)(row - (int64_t)node * B);
      v = xpin[(int64_t)k * xp_k + (int64_t)node * xp_n + (int64_t)b * cin + ci];
      if (rnd) v = tf32_rn(v);
    } else if (j == NB * cin) {
      v = 1.0f;
    }
    if (ib16 != nullptr) {                 // fp16 fused path: the tensor-core operand is the half copy
      ib16[i] = __float2half_rn(v);
      v = __half2float(__float2half_rn(v));
      if (ib_g == nullptr) continue;       // eval: the fp32 copies are only read by the backward
    }
    ib_g[i] = v;
    ib_u[i] = v;
  }
}

// d(input block) [R][hs] (columns k*cin+ci) of the update and gate AGCNs -> dXPin [NB][R][cin] (their sum).
__global__ void k_repack_dib(const float* __restrict__ dib_a, const float* __restrict__ dib_b, int NB, int cin,
                             int64_t R, int hs, float* __restrict__ dxpin, int rnd) {
  const int64_t total = (int64_t)NB * R * cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % cin);
    const int64_t row = (i / cin) % R;
    const int k = (int)(i / ((int64_t)cin * R));
    const int64_t src = row * hs + k * cin + ci;
    float v = dib_a[src] + dib_b[src];
    dxpin[i] = rnd ? tf32_rn(v) : v;
  }
}

// ---- input staging ------------------------------------------------------------------
// Encoder inputs for all steps: x [B][T][N][Cin] -> XPin block 0, layout [N][T][B][Cin].
__global__ void k_stage_encoder_input(const float* __restrict__ x, float* __restrict__ xp0, int B, int T, int N, int Cin,
                                      int rnd) {
  int64_t total = (int64_t)N * T * B * Cin;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int ci = (int)(i % Cin);
    int b = (int)((i / Cin) % B);
    int t = (int)((i / ((int64_t)Cin * B)) % T);
    int n = (int)(i / ((int64_t)Cin * B * T));
    float v = x[(((int64_t)b * T + t) * N + n) * Cin + ci];
    xp0[i] = rnd ? tf32_rn(v) : v;
  }
}

// Decoder input of step t: [go | y_cov[:,t]] -> XPin block 0 [N][B][Cout+Ycov]  (model/MegaCRN.py:185).
// go_src: nullptr (t == 0, zeros), or a [B][T][N][Cout] tensor (output or labels) read at step t-1.
__global__ void k_stage_decoder_input(const float* __restrict__ go_src, const float* __restrict__ ycov,
                                      float* __restrict__ xp0, int B, int T, int N, int Cout, int Ycov, int t, int rnd) {
  int C = Cout + Ycov;
  int64_t total = (int64_t)N * B * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int b = (int)((i / C) % B);
    int n = (int)(i / ((int64_t)C * B));
    float v;
    if (c < Cout) v = go_src ? go_src[(((int64_t)b * T + (t - 1)) * N + n) * Cout + c] : 0.f;
    else v = ycov[(((int64_t)b * T + t) * N + n) * Ycov + (c - Cout)];
    xp0[i] = rnd ? tf32_rn(v) : v;
  }
}

// ---- projection (model/MegaCRN.py:186): out[b][t][n][co] = h[n][b][:] . wp[co][:] + bp[co]; warp per row
__global__ void __launch_bounds__(256) k_proj_fwd(const float* __restrict__ h, const float* __restrict__ wp,
                                                  const float* __restrict__ bp, float* __restrict__ out,
                                                  int B, int T, int N, int D, int Cout, int t) {
  int64_t row = blockIdx.x * 8 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= (int64_t)N * B) return;
  int n = (int)(row / B), b = (int)(row % B);
  const float* hr = h + row * D;
  for (int co = 0; co < Cout; ++co) {
    float s = 0.f;
    for (int j = lane; j < D; j += 32) s = fmaf(hr[j], wp[(int64_t)co * D + j], s);
    s = warp_sum(s);
    if (lane == 0) out[(((int64_t)b * T + t) * N + n) * Cout + co] = s + bp[co];
  }
}

// Projection of several decoder steps in one launch (steps = set bits of tmask; blockIdx.y selects the y-th one):
// h_t = hx_base + (t + 1) * hx_step for t + 1 < T, else h_last (the per-step states kept for the backward).
__global__ void __launch_bounds__(256) k_proj_fwd_steps(const float* __restrict__ hx_base, int64_t hx_step,
                                                        const float* __restrict__ h_last, const float* __restrict__ wp,
                                                        const float* __restrict__ bp, float* __restrict__ out, unsigned tmask,
                                                        int B, int T, int N, int D, int Cout) {
  int t = 0;
  {
    unsigned m = tmask;
    for (int y = blockIdx.y; y > 0; --y) m &= m - 1;
    t = __ffs(m) - 1;
  }
  const float* h = (t + 1 < T) ? hx_base + (int64_t)(t + 1) * hx_step : h_last;
  int64_t row = blockIdx.x * 8 + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (row >= (int64_t)N * B) return;
  int n = (int)(row / B), b = (int)(row % B);
  const float* hr = h + row * D;
  for (int co = 0; co < Cout; ++co) {
    float s = 0.f;
    for (int j = lane; j < D; j += 32) s = fmaf(hr[j], wp[(int64_t)co * D + j], s);
    s = warp_sum(s);
    if (lane == 0) out[(((int64_t)b * T + t) * N + n) * Cout + co] = s + bp[co];
  }
}

// Projection backward for step t.  d_out_t[n][b][co] = dOut[b][t][n][co] (+ dgo[n][b][co] when the next
// decoder input was this step's own prediction); dH[r][:] (+)= d_out_t[r] . wp; dwp, dbp accumulate.
// One block = 32 rows; thread j owns column(s) j of D.
__global__ void __launch_bounds__(256) k_proj_bwd(const float* __restrict__ dOut, const float* __restrict__ dxin,
                                                  int dxin_stride, const float* __restrict__ h,
                                                  const float* __restrict__ wp, float* __restrict__ dH, int dh_init,
                                                  float* __restrict__ dwp, float* __restrict__ dbp,
                                                  int B, int T, int N, int D, int Cout, int t) {
  extern __shared__ float sh_do[];                 // [32][Cout]
  int64_t R = (int64_t)N * B, r0 = (int64_t)blockIdx.x * 32;
  for (int i = threadIdx.x; i < 32 * Cout; i += blockDim.x) {
    int64_t row = r0 + i / Cout;
    int co = i % Cout;
    float v = 0.f;
    if (row < R) {
      int n = (int)(row / B), b = (int)(row % B);
      if (dOut) v = dOut[(((int64_t)b * T + t) * N + n) * Cout + co];
      if (dxin) v += dxin[row * dxin_stride + co];
    }
    sh_do[i] = v;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    for (int co = 0; co < Cout; ++co) {
      float accw = 0.f;
      for (int i = 0; i < 32; ++i) {
        int64_t row = r0 + i;
        if (row < R) accw = fmaf(sh_do[i * Cout + co], h[row * D + j], accw);
      }
      atomicAdd(dwp + (int64_t)co * D + j, accw);
    }
    for (int i = 0; i < 32; ++i) {
      int64_t row = r0 + i;
      if (row < R) {
        float v = dh_init ? 0.f : dH[row * D + j];
        for (int co = 0; co < Cout; ++co) v = fmaf(sh_do[i * Cout + co], wp[(int64_t)co * D + j], v);
        dH[row * D + j] = v;
      }
    }
  }
  if (threadIdx.x < Cout) {
    float s = 0.f;
    for (int i = 0; i < 32; ++i) s += sh_do[i * Cout + threadIdx.x];
    atomicAdd(dbp + threadIdx.x, s);
  }
}

// ---- memory-bank query (model/MegaCRN.py:159-166, :179); one warp per (node, batch) row ----
// h [R][H] node-major.  Writes query/value node-major [R][d] (for backward), att [R][M], ind [R][2],
// the four batch-major outputs [B][N][d], and the decoder's initial state [R][H+d] = [h | value].
constexpr int MQ_ROWS_PER_WARP = 1;               // (a 64-rows-per-block variant with Wq staged in shared memory measured slower: 70 vs 53 us)
__global__ void __launch_bounds__(256) k_memory_query(const float* __restrict__ h, const float* __restrict__ wq,
                                                      const float* __restrict__ mem, float* __restrict__ q_nm,
                                                      float* __restrict__ att, int* __restrict__ ind,
                                                      float* __restrict__ o_hatt, float* __restrict__ o_query,
                                                      float* __restrict__ o_pos, float* __restrict__ o_neg,
                                                      float* __restrict__ dec_h0, float* __restrict__ dec_h0_mma,
                                                      __half* __restrict__ dec_x16, int rnd, int B, int N, int H, int M, int d) {
  // shared: per warp h row [H] + q[d] + sc[M]; the memory bank [M][d + 1] (padded: lanes that walk different memory rows hit
  // different banks)
  extern __shared__ float shm[];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = H + d + M, dp = d + 1;
  float* memS = shm + 8 * per_warp;
  for (int i = threadIdx.x; i < M * d; i += blockDim.x) memS[(i / d) * dp + (i % d)] = mem[i];
  __syncthreads();
  float* hs = shm + warp * per_warp;
  float* q = hs + H;
  float* sc = q + d;
  const int64_t R = (int64_t)N * B;
  for (int rr = 0; rr < MQ_ROWS_PER_WARP; ++rr) {
    const int64_t row = ((int64_t)blockIdx.x * 8 + warp) * MQ_ROWS_PER_WARP + rr;
    if (row >= R) break;
    int n = (int)(row / B), b = (int)(row % B);
    const float* hr = h + row * H;
    __syncwarp();
    for (int j = lane; j < H; j += 32) hs[j] = hr[j];
    __syncwarp();
    for (int j = lane; j < d; j += 32) {             // query = h Wq            :160
      float s = 0.f;
      for (int k = 0; k < H; ++k) s = fmaf(hs[k], wq[(int64_t)k * d + j], s);
      q[j] = s;
    }
    __syncwarp();
    for (int m = lane; m < M; m += 32) {             // logits = q Mem^T        :161
      float s = 0.f;
      for (int k = 0; k < d; ++k) s = fmaf(q[k], memS[m * dp + k], s);
      sc[m] = s;
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int m = lane; m < M; m += 32) mx = fmaxf(mx, sc[m]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int m = lane; m < M; m += 32) sum += expf(sc[m] - mx);
    sum = warp_sum(sum);
    float inv = 1.0f / sum;
    __syncwarp();
    for (int m = lane; m < M; m += 32) sc[m] = expf(sc[m] - mx) * inv;
    __syncwarp();
    // top-2 (first index wins ties, as torch.topk does on sorted-descending stable order)   :163
    int i0 = 0, i1 = -1;
    if (lane == 0) {
      float b0 = sc[0], b1 = -INFINITY;
      for (int m = 1; m < M; ++m) {
        float v = sc[m];
        if (v > b0) { b1 = b0; i1 = i0; b0 = v; i0 = m; }
        else if (v > b1) { b1 = v; i1 = m; }
      }
      if (i1 < 0) i1 = 0;
    }
    i0 = __shfl_sync(0xffffffffu, i0, 0);
    i1 = __shfl_sync(0xffffffffu, i1, 0);
    int64_t ob = ((int64_t)b * N + n) * d;
    const int64_t o0 = row * (H + d);
    for (int j = lane; j < d; j += 32) {             // value = att Mem          :162
      float s = 0.f;
      for (int m = 0; m < M; ++m) s = fmaf(sc[m], memS[m * dp + j], s);
      o_hatt[ob + j] = s;
      o_query[ob + j] = q[j];
      o_pos[ob + j] = memS[i0 * dp + j];             // :164
      o_neg[ob + j] = memS[i1 * dp + j];             // :165
      dec_h0[o0 + H + j] = s;                        // :179
      dec_h0_mma[o0 + H + j] = rnd ? tf32_rn(s) : s;
      if (dec_x16) dec_x16[o0 + H + j] = __float2half_rn(s);      // fp16 operand copy of the decoder's initial state
      if (q_nm) q_nm[row * d + j] = q[j];
    }
    for (int j = lane; j < H; j += 32) {
      dec_h0[o0 + j] = hs[j];
      dec_h0_mma[o0 + j] = rnd ? tf32_rn(hs[j]) : hs[j];
      if (dec_x16) dec_x16[o0 + j] = __float2half_rn(hs[j]);
    }
    if (att) for (int m = lane; m < M; m += 32) att[row * M + m] = sc[m];
    if (ind && lane == 0) { ind[row * 2] = i0; ind[row * 2 + 1] = i1; }
  }
}

// Row part of the memory-query backward (tests/kernel_spec.py:memory_query_bwd).  One warp per row.
//  d_value[r] = dH0[r][H:] + d_hatt ; d_att = d_value Mem^T ; d_sc = att*(d_att - sum(att*d_att)) ;
//  d_q = d_query + d_sc Mem.  Writes d_value, d_sc, d_q (node-major) for the GEMM reductions that follow;
//  scatters d_pos/d_neg into dMem with atomics (these are null when the trainer detaches them).
__global__ void __launch_bounds__(256) k_memory_query_bwd_rows(
    const float* __restrict__ dH0, const float* __restrict__ d_hatt, const float* __restrict__ d_query,
    const float* __restrict__ d_pos, const float* __restrict__ d_neg, const float* __restrict__ mem,
    const float* __restrict__ att, const int* __restrict__ ind, float* __restrict__ dv, float* __restrict__ dsc,
    float* __restrict__ dq, float* __restrict__ dMem, int B, int N, int H, int M, int d) {
  extern __shared__ float shm[];                   // per warp: dv[d] + ds[M]; then the memory bank [M][d + 1]
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* memS = shm + 8 * (d + M);
  const int dp = d + 1;
  for (int i = threadIdx.x; i < M * d; i += blockDim.x) memS[(i / d) * dp + (i % d)] = mem[i];
  __syncthreads();
  int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= (int64_t)N * B) return;
  float* v = shm + warp * (d + M);
  float* ds = v + d;
  int n = (int)(row / B), b = (int)(row % B);
  int64_t ob = ((int64_t)b * N + n) * d;
  for (int j = lane; j < d; j += 32) {
    float t = dH0[row * (H + d) + H + j];
    if (d_hatt) t += d_hatt[ob + j];
    v[j] = t;
    dv[row * d + j] = t;
  }
  __syncwarp();
  float part = 0.f;
  for (int m = lane; m < M; m += 32) {
    float s = 0.f;
    for (int k = 0; k < d; ++k) s = fmaf(v[k], memS[m * dp + k], s);
    ds[m] = s;                                     // d_att
    part = fmaf(att[row * M + m], s, part);
  }
  part = warp_sum(part);
  __syncwarp();
  for (int m = lane; m < M; m += 32) {
    float t = att[row * M + m] * (ds[m] - part);
    ds[m] = t;
    dsc[row * M + m] = t;
  }
  __syncwarp();
  int i0 = ind[row * 2], i1 = ind[row * 2 + 1];
  for (int j = lane; j < d; j += 32) {
    float s = d_query ? d_query[ob + j] : 0.f;
    for (int m = 0; m < M; ++m) s = fmaf(ds[m], memS[m * dp + j], s);
    dq[row * d + j] = s;
    if (d_pos) atomicAdd(dMem + (int64_t)i0 * d + j, d_pos[ob + j]);
    if (d_neg) atomicAdd(dMem + (int64_t)i1 * d + j, d_neg[ob + j]);
  }
}

// ---- cell backward, elementwise pieces (tests/kernel_spec.py:cell_bwd) ----------------------
// dU = dH' * (1-r) * (1-hc^2)
__global__ void k_bwd_du(const float* __restrict__ dH, const float* __restrict__ r, const float* __restrict__ hc,
                         float* __restrict__ dU, int64_t n, int rnd) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float c = hc[i];
    float v = dH[i] * (1.0f - r[i]) * (1.0f - c * c);
    dU[i] = rnd ? tf32_rn(v) : v;
  }
}

// a += alpha * b
__global__ void k_axpy(float* __restrict__ a, const float* __restrict__ b, float alpha, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    a[i] = fmaf(alpha, b[i], a[i]);
}

__global__ void k_add_inplace(float* __restrict__ a, const float* __restrict__ b, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    a[i] += b[i];
}

// out[r][0:H] = a[r][0:H] (row stride lda) + b[r][0:H]
__global__ void k_slice_add(const float* __restrict__ a, int lda, const float* __restrict__ b, float* __restrict__ out,
                            int64_t R, int H) {
  int64_t n = R * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / H;
    int c = (int)(i - row * H);
    out[i] = a[row * lda + c] + b[i];
  }
}

// ---- stacked cells (num_layers > 1; tests/kernel_spec.py: cell_fwd_wide / cell_bwd_wide) ----------------------
// AGCN operand of a stacked cell: out[row][0:hs] = a[row] (the state of the layer below), out[row][hs:2hs] = b[row]
// (own state h, or z*h for the candidate: model/MegaCRN.py:42, :46); TF32-rounded when rnd (tensor-core operand).
__global__ void k_concat2(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int64_t R,
                          int hs, int rnd) {
  const int kw = 2 * hs;
  const int64_t total = R * kw;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / kw;
    const int c = (int)(i - row * kw);
    const float v = (c < hs) ? a[row * hs + c] : b[row * hs + (c - hs)];
    out[i] = rnd ? tf32_rn(v) : v;
  }
}

// Gate backward of a stacked cell.  dV0 [R][2hs] = gradient of the candidate operand [x_in | z*h]:
//   dG[:, :hs] = dZH*h*z(1-z), dG[:, hs:] = dH*(h-hc)*r(1-r), dh_part = dH*r + dZH*z, dx_part = dV0[:, :hs].
__global__ void k_wide_gate_bwd(const float* __restrict__ dV0, const float* __restrict__ dH, const float* __restrict__ h,
                                const float* __restrict__ z, const float* __restrict__ r, const float* __restrict__ hc,
                                float* __restrict__ dG, float* __restrict__ dh_part, float* __restrict__ dx_part, int64_t R,
                                int hs, int rnd) {
  const int64_t total = R * hs;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / hs;
    const int c = (int)(i - row * hs);
    const float dzh = dV0[row * 2 * hs + hs + c], hh = h[i], zz = z[i], rr = r[i], dh = dH[i];
    float gz = dzh * hh * zz * (1.0f - zz);
    float gr = dh * (hh - hc[i]) * rr * (1.0f - rr);
    if (rnd) { gz = tf32_rn(gz); gr = tf32_rn(gr); }
    dG[row * 2 * hs + c] = gz;
    dG[row * 2 * hs + hs + c] = gr;
    dh_part[i] = dh * rr + dzh * zz;
    dx_part[i] = dV0[row * 2 * hs + c];
  }
}

// dV0 [R][2hs] = gradient of the gate operand [x_in | h]:  dH = dh_part + dV0[:, hs:],  dx = dx_part + dV0[:, :hs].
__global__ void k_wide_finish(const float* __restrict__ dV0, const float* __restrict__ dh_part,
                              const float* __restrict__ dx_part, float* __restrict__ dH, float* __restrict__ dx, int64_t R,
                              int hs) {
  const int64_t total = R * hs;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / hs;
    const int c = (int)(i - row * hs);
    dH[i] = dh_part[i] + dV0[row * 2 * hs + hs + c];
    dx[i] = dx_part[i] + dV0[row * 2 * hs + c];
  }
}

}  // namespace mcrn
