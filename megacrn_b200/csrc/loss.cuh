// The training step's loss, fused: masked MAE on inverse-scaled values
// (model/utils.py:126-133, :45-54) + lamb*TripletMarginLoss(margin=1) + lamb1*MSELoss on
// query vs. (detached) pos/neg (model/traintest_MegaCRN.py:118-125).
#pragma once

#include "small_kernels.cuh"

namespace mcrn {

// scratch[0] = #(y_true != 0), [1] = sum |y_pred-y_true|*mask, [2] = sum triplet hinge, [3] = sum (q-p)^2
__global__ void __launch_bounds__(256) k_loss_reduce_out(const float* __restrict__ out, const float* __restrict__ lab,
                                                         int64_t n, float mean, float std, float* __restrict__ scratch) {
  __shared__ float sh[8];
  float cnt = 0.f, sabs = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float yt = __fadd_rn(__fmul_rn(lab[i], std), mean);
    float yp = __fadd_rn(__fmul_rn(out[i], std), mean);
    if (yt != 0.f) { cnt += 1.f; sabs += fabsf(yp - yt); }
  }
  cnt = block_sum_256(cnt, sh);
  sabs = block_sum_256(sabs, sh);
  if (threadIdx.x == 0) { atomicAdd(scratch + 0, cnt); atomicAdd(scratch + 1, sabs); }
}

// one warp per (b, n) row of query/pos/neg [rows][d]
__global__ void __launch_bounds__(256) k_loss_reduce_rows(const float* __restrict__ q, const float* __restrict__ p,
                                                          const float* __restrict__ ng, int64_t rows, int d,
                                                          float* __restrict__ scratch) {
  __shared__ float sh[8];
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float trip = 0.f, mse = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < rows; row += (int64_t)gridDim.x * 8) {
    float sp = 0.f, sn = 0.f, sm = 0.f;
    for (int j = lane; j < d; j += 32) {
      float qq = q[row * d + j], pp = p[row * d + j], nn = ng[row * d + j];
      float a = qq - pp + 1e-6f, b = qq - nn + 1e-6f, c = qq - pp;
      sp = fmaf(a, a, sp); sn = fmaf(b, b, sn); sm = fmaf(c, c, sm);
    }
    sp = warp_sum(sp); sn = warp_sum(sn); sm = warp_sum(sm);
    if (lane == 0) { trip += fmaxf(sqrtf(sp) - sqrtf(sn) + 1.0f, 0.f); mse += sm; }
  }
  trip = block_sum_256(trip, sh);
  mse = block_sum_256(mse, sh);
  if (threadIdx.x == 0) { atomicAdd(scratch + 2, trip); atomicAdd(scratch + 3, mse); }
}

// this rank's count of labels whose inverse-scaled value is non-zero, added to out[0] (model/utils.py:127)
__global__ void __launch_bounds__(256) k_mask_count(const float* __restrict__ lab, int64_t n, float mean, float std, float* __restrict__ out) {
  __shared__ float sh[8];
  float cnt = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (__fadd_rn(__fmul_rn(lab[i], std), mean) != 0.f) cnt += 1.f;
  cnt = block_sum_256(cnt, sh);
  if (threadIdx.x == 0) atomicAdd(out, cnt);
}

// cnt_override (device, or null): the masked-MAE normaliser to use instead of this batch's own count (data parallel:
// global count / world size)
__global__ void k_loss_finish(const float* __restrict__ scratch, const float* __restrict__ cnt_override, int64_t rows, int d, float lamb,
                              float lamb1, float* __restrict__ loss_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float cnt = cnt_override ? cnt_override[0] : scratch[0];
    float l1 = cnt > 0.f ? scratch[1] / cnt : 0.f;
    loss_out[0] = l1 + lamb * scratch[2] / (float)rows + lamb1 * scratch[3] / ((float)rows * (float)d);
  }
}

__global__ void k_loss_grad_out(const float* __restrict__ out, const float* __restrict__ lab, int64_t n, float mean,
                                float std, const float* __restrict__ scratch, const float* __restrict__ cnt_override,
                                float* __restrict__ d_out) {
  float cnt = cnt_override ? cnt_override[0] : scratch[0];
  float sc = cnt > 0.f ? std / cnt : 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float yt = __fadd_rn(__fmul_rn(lab[i], std), mean);
    float yp = __fadd_rn(__fmul_rn(out[i], std), mean);
    float df = yp - yt;
    float g = (yt != 0.f) ? (df > 0.f ? sc : (df < 0.f ? -sc : 0.f)) : 0.f;
    d_out[i] = g;
  }
}

__global__ void __launch_bounds__(256) k_loss_grad_rows(const float* __restrict__ q, const float* __restrict__ p,
                                                        const float* __restrict__ ng, int64_t rows, int d, float lamb,
                                                        float lamb1, float* __restrict__ d_q) {
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < rows; row += (int64_t)gridDim.x * 8) {
    float sp = 0.f, sn = 0.f;
    for (int j = lane; j < d; j += 32) {
      float qq = q[row * d + j];
      float a = qq - p[row * d + j] + 1e-6f, b = qq - ng[row * d + j] + 1e-6f;
      sp = fmaf(a, a, sp); sn = fmaf(b, b, sn);
    }
    sp = sqrtf(warp_sum(sp)); sn = sqrtf(warp_sum(sn));
    bool active = (sp - sn + 1.0f) > 0.f;
    float ct = lamb / (float)rows, cm = 2.0f * lamb1 / ((float)rows * (float)d);
    for (int j = lane; j < d; j += 32) {
      float qq = q[row * d + j], pp = p[row * d + j], nn = ng[row * d + j];
      float g = cm * (qq - pp);
      if (active) g += ct * ((qq - pp + 1e-6f) / sp - (qq - nn + 1e-6f) / sn);
      d_q[row * d + j] = g;
    }
  }
}

}  // namespace mcrn

// ---- fused gradient clipping + Adam over the 14 parameter tensors (model/traintest_MegaCRN.py:104, :129-130) ----
namespace mcrn {
// NT = 14 (num_layers == 1) or 14 + 8*(MCRN_MAX_LAYERS-1) (stacked cells; unused trailing entries have zero elements)
template <int NT>
struct ParamTableT {
  float* p[NT]; const float* g[NT]; float* m[NT]; float* v[NT];
  int64_t off[NT + 1];                   // prefix sums of the element counts
};
using ParamTable = ParamTableT<14>;
constexpr int PARAM_TABLE_MAX = 14 + 8 * (MCRN_MAX_LAYERS - 1);
template <int NT>
__device__ __forceinline__ int table_find(const ParamTableT<NT>& t, int64_t i) {
  int k = 0;
#pragma unroll
  for (int j = 1; j < NT; ++j) k += (i >= t.off[j]) ? 1 : 0;
  return k;
}
// state[0] = step count (incremented here), state[2] += sum of squared gradients (zeroed by the caller)
template <int NT>
__global__ void __launch_bounds__(256) k_grad_sqnorm(ParamTableT<NT> t, float* __restrict__ state) {
  __shared__ float sh[8];
  float s = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < t.off[NT]; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = table_find(t, i);
    const float g = t.g[k][i - t.off[k]];
    s = fmaf(g, g, s);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    s = sh[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffu, s, o);
    if (threadIdx.x == 0) atomicAdd(state + 2, s);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) state[0] += 1.0f;
}
// torch.nn.utils.clip_grad_norm_ (coef = min(1, max_norm / (norm + 1e-6))) then torch.optim.Adam.step (no amsgrad, no decay)
template <int NT>
__global__ void __launch_bounds__(256) k_clip_adam(ParamTableT<NT> t, float* __restrict__ state, float beta1, float beta2, float eps,
                                                   float max_norm) {
  const float step = state[0], lr = state[1];
  const float norm = sqrtf(state[2]);
  // a non-finite gradient norm (fp16 operand overflow upstream, exploding gradients) must not reach the weights: the step
  // is skipped (clip coefficient reported as 0; the step counter has already advanced, as torch's GradScaler-style skips do not)
  if (!(norm < INFINITY)) {
    if (blockIdx.x == 0 && threadIdx.x == 0) state[3] = 0.0f;
    return;
  }
  const float coef = max_norm > 0.f ? fminf(1.0f, max_norm / (norm + 1e-6f)) : 1.0f;
  const float bc1 = 1.0f - powf(beta1, step), bc2 = 1.0f - powf(beta2, step);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < t.off[NT]; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = table_find(t, i);
    const int64_t j = i - t.off[k];
    const float g = t.g[k][j] * coef;
    const float m = beta1 * t.m[k][j] + (1.0f - beta1) * g;
    const float v = beta2 * t.v[k][j] + (1.0f - beta2) * g * g;
    t.m[k][j] = m;
    t.v[k][j] = v;
    t.p[k][j] -= step_size * m / (sqrtf(v) * inv_sqrt_bc2 + eps);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) state[3] = coef;
}
}  // namespace mcrn
