// GEMM engine interface shared by the SIMT fp32 path (gemm_simt.cuh) and the
// tcgen05 TF32 path (gemm_tc.cuh): one logical description of the contraction and
// a family of epilogue functors that fuse the surrounding elementwise work.
//
//   C[batch][m][n] = sum_{seg<nseg} sum_{k<Kseg} A(batch, m, seg, k) * B(batch, seg, k, n)
//
// Operands are addressed through element strides, so K-contiguous ("K-major") and
// M/N-contiguous ("MN-major") storage, block-segmented K (the [X | P_1..P_KS] buffers)
// and batching are all the same kernel.
#pragma once

#include "common.cuh"

namespace mcrn {

struct GemmDesc {
  const float* A = nullptr;
  int64_t a_row = 0, a_k = 0, a_seg = 0, a_batch = 0;
  const float* B = nullptr;
  int64_t b_k = 0, b_n = 0, b_seg = 0, b_batch = 0;
  int M = 0, N = 0, Kseg = 0, nseg = 1, nbatch = 1, splits = 1;
};

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- epilogues ------------------------------------------------------------------
// Protocol: op.template apply<V>(batch, m, n0, nv, acc) is called once per output row
// fragment of V contiguous columns starting at n0 (nv <= V of them in range).

// C = alpha*acc (+ add1 + add2), row-major with leading dimension ldc.
struct EpiStore {
  float* C; int64_t ldc, c_batch; float alpha;
  const float* add1; const float* add2;     // optional, same layout as C
  template <int V>
  __device__ __forceinline__ void apply(int bz, int m, int n0, int nv, const float (&acc)[V]) const {
    int64_t off = (int64_t)bz * c_batch + (int64_t)m * ldc + n0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (j < nv) {
        float v = alpha * acc[j];
        if (add1) v += add1[off + j];
        if (add2) v += add2[off + j];
        C[off + j] = v;
      }
    }
  }
};

// C[m][n] = acc + add[m*ld_add + n]   (add has a different row stride, e.g. a column slice)
struct EpiStoreStrideAdd {
  float* C; int64_t ldc; const float* add; int64_t ld_add;
  template <int V>
  __device__ __forceinline__ void apply(int, int m, int n0, int nv, const float (&acc)[V]) const {
#pragma unroll
    for (int j = 0; j < V; ++j)
      if (j < nv) C[(int64_t)m * ldc + n0 + j] = acc[j] + add[(int64_t)m * ld_add + n0 + j];
  }
};

// C += acc with atomics (split-K / accumulation over timesteps).
struct EpiAtomicAdd {
  float* C; int64_t ldc, c_batch;
  template <int V>
  __device__ __forceinline__ void apply(int bz, int m, int n0, int nv, const float (&acc)[V]) const {
    int64_t off = (int64_t)bz * c_batch + (int64_t)m * ldc + n0;
#pragma unroll
    for (int j = 0; j < V; ++j)
      if (j < nv) atomicAdd(C + off + j, acc[j]);
  }
};

// Chebyshev recursion  T_k = 2*(S*T_{k-1}) - T_{k-2}  (model/MegaCRN.py:21-22);
// prev == nullptr means T_{k-2} = I.
struct EpiCheb {
  float* C; int64_t ldc; const float* prev;
  template <int V>
  __device__ __forceinline__ void apply(int, int m, int n0, int nv, const float (&acc)[V]) const {
    int64_t off = (int64_t)m * ldc + n0;
#pragma unroll
    for (int j = 0; j < V; ++j)
      if (j < nv) C[off + j] = 2.0f * acc[j] - (prev ? prev[off + j] : ((n0 + j) == m ? 1.0f : 0.0f));
  }
};

// Column n = blk*W + c is stored at C[blk][m][c]  (the dXP block buffers).
struct EpiBlocks {
  float* C; int W; int64_t blk_stride;
  template <int V>
  __device__ __forceinline__ void apply(int, int m, int n0, int nv, const float (&acc)[V]) const {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (j < nv) {
        int n = n0 + j, blk = n / W, c = n - blk * W;
        C[(int64_t)blk * blk_stride + (int64_t)m * W + c] = acc[j];
      }
    }
  }
};

// Input-channel contribution of an AGCN (the 1-2 raw input channels):
//   sum_{k<NB} sum_{ci<Cin} XPin[k][node][.][b][ci] * Win[k][ci][col]  + bias[col]
struct InTerm {
  const float* xpin; int64_t xp_k, xp_n;   // XPin(k, node, b, ci) = xpin + k*xp_k + node*xp_n + b*Cin + ci
  const float* win;                        // [NB][Cin][O]
  const float* bias;                       // [O]
  int NB, Cin, O, Bsz;
  __device__ __forceinline__ float eval(int m, int col) const {
    int node = m / Bsz, b = m - node * Bsz;
    const float* xp = xpin + (int64_t)node * xp_n + (int64_t)b * Cin;
    float s = bias[col];
    for (int k = 0; k < NB; ++k)
      for (int ci = 0; ci < Cin; ++ci)
        s = fmaf(xp[k * xp_k + ci], win[((int64_t)k * Cin + ci) * O + col], s);
    return s;
  }
};

// Gate AGCN epilogue (model/MegaCRN.py:43-45): zr = sigmoid(acc + in-term + bias);
// columns [0,H) are z, [H,2H) are r;  writes z, r and z*h (the update AGCN's state input).
struct EpiGate {
  InTerm in; int H;
  const float* h;   // [R][H] current state (XPg block 0)
  float* z; float* r; float* zh;
  template <int V>
  __device__ __forceinline__ void apply(int, int m, int n0, int nv, const float (&acc)[V]) const {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (j < nv) {
        int col = n0 + j;
        float s = sigmoid_f(acc[j] + in.eval(m, col));
        if (col < H) {
          int64_t o = (int64_t)m * H + col;
          if (z) z[o] = s;
          zh[o] = s * h[o];
        } else {
          r[(int64_t)m * H + (col - H)] = s;
        }
      }
    }
  }
};

// Update AGCN epilogue (model/MegaCRN.py:46-47): hc = tanh(acc + in-term + bias);
// h' = r*h + (1-r)*hc.
struct EpiUpdate {
  InTerm in; int H;
  const float* h; const float* r;
  float* hc; float* h_out;
  template <int V>
  __device__ __forceinline__ void apply(int, int m, int n0, int nv, const float (&acc)[V]) const {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (j < nv) {
        int col = n0 + j;
        int64_t o = (int64_t)m * H + col;
        float c = tanhf(acc[j] + in.eval(m, col));
        float rr = r[o];
        if (hc) hc[o] = c;
        h_out[o] = rr * h[o] + (1.0f - rr) * c;
      }
    }
  }
};

}  // namespace mcrn
