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
  int a_nseg = 0, b_nseg = 0;   // distinct segments of A / B (0 = nseg); segment index wraps (hi/lo weight split)
  int prec_exact = 0;           // 1 = this contraction must run in exact fp32 (SIMT engine)
  int use_map = 0;              // 1 = K-segment s reads A segment a_map[s] and B segment b_map[s] (3xTF32 products)
  uint8_t a_map[16] = {0}, b_map[16] = {0};
  __host__ __device__ int nseg_a() const { return a_nseg ? a_nseg : nseg; }
  __host__ __device__ int nseg_b() const { return b_nseg ? b_nseg : nseg; }
  __host__ __device__ int seg_a(int s) const { return use_map ? a_map[s] : s % nseg_a(); }
  __host__ __device__ int seg_b(int s) const { return use_map ? b_map[s] : s % nseg_b(); }
};

// Round-to-nearest TF32 (cvt.rna): producers of tensor-core operands store rounded values so that the
// tensor core's mantissa truncation is exact and the contraction behaves as round-to-nearest TF32.
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- epilogues ------------------------------------------------------------------
// Protocol: op.template apply<V>(batch, m, n0, nv, acc) is called once per output row
// fragment of V contiguous columns starting at n0 (nv <= V of them in range).

// C = alpha*acc (+ add1 + add2), row-major with leading dimension ldc.
struct EpiStore {
  float* C; int64_t ldc, c_batch; float alpha;
  const float* add1; const float* add2;     // optional, same layout as C
  int rnd = 0;                              // 1: store TF32-rounded (the output is a tensor-core operand)
  template <int V>
  __device__ __forceinline__ void apply(int bz, int m, int n0, int nv, const float (&acc)[V]) const {
    int64_t off = (int64_t)bz * c_batch + (int64_t)m * ldc + n0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (j < nv) {
        float v = alpha * acc[j];
        // ld.global.nc for addends that are not the output itself: lets the compiler batch the loads of the
        // unrolled row loop instead of serialising load -> store -> load on a possible alias
        if (add1) v += (add1 != C) ? __ldg(add1 + off + j) : add1[off + j];
        if (add2) v += (add2 != C) ? __ldg(add2 + off + j) : add2[off + j];
        C[off + j] = rnd ? tf32_rn(v) : v;
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
  float* Cr;                                 // TF32-rounded copy (tensor-core operand), may alias nothing
  template <int V>
  __device__ __forceinline__ void apply(int, int m, int n0, int nv, const float (&acc)[V]) const {
    int64_t off = (int64_t)m * ldc + n0;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (j < nv) {
        float v = 2.0f * acc[j] - (prev ? prev[off + j] : ((n0 + j) == m ? 1.0f : 0.0f));
        C[off + j] = v;
        if (Cr) Cr[off + j] = tf32_rn(v);
      }
    }
  }
};

// Column n = blk*W + c is stored at C[blk][m][c]  (the dXP block buffers).
struct EpiBlocks {
  float* C; int W; int64_t blk_stride;
  int rnd = 0;                               // 1: blocks 1..last-1 (tensor-core operands downstream) are TF32-rounded
  int last = -1;                             // index of the input block (never rounded)
  float* last_out = nullptr;                 // if set, block `last` is written here ([M][W]) instead of C[last]
  template <int V>
  __device__ __forceinline__ void apply(int, int m, int n0, int nv, const float (&acc)[V]) const {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (j < nv) {
        int n = n0 + j, blk = n / W, c = n - blk * W;
        if (blk == last && last_out) {
          last_out[(int64_t)m * W + c] = acc[j];
        } else {
          int64_t o = (int64_t)blk * blk_stride + (int64_t)m * W + c;
          C[o] = (rnd && blk > 0 && blk != last) ? tf32_rn(acc[j]) : acc[j];
        }
      }
    }
  }
};

// Gate AGCN epilogue (model/MegaCRN.py:43-45): zr = sigmoid(acc)  (input channels and bias are the
// "input block" of the contraction); columns [0,H) are z, [H,2H) are r; writes z, r and z*h.
struct EpiGate {
  int H;
  const float* h;   // [R][H] current state, exact fp32
  float* z; float* r; float* zh;
  int rnd;          // 1: zh (a tensor-core operand only) is stored TF32-rounded
  template <int V>
  __device__ __forceinline__ void apply(int, int m, int n0, int nv, const float (&acc)[V]) const {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (j < nv) {
        int col = n0 + j;
        float s = sigmoid_f(acc[j]);
        if (col < H) {
          int64_t o = (int64_t)m * H + col;
          if (z) z[o] = s;
          float t = s * __ldg(h + o);
          zh[o] = rnd ? tf32_rn(t) : t;
        } else {
          r[(int64_t)m * H + (col - H)] = s;
        }
      }
    }
  }
};

// Update AGCN epilogue (model/MegaCRN.py:46-47): hc = tanh(acc); h' = r*h + (1-r)*hc.
struct EpiUpdate {
  int H;
  const float* h; const float* r;   // exact state, r gate
  float* hc; float* h_out;          // h_out: exact new state
  float* h_mma; int rnd;            // h_mma: the copy the next propagation / gate GEMM reads (TF32-rounded if rnd)
  template <int V>
  __device__ __forceinline__ void apply(int, int m, int n0, int nv, const float (&acc)[V]) const {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (j < nv) {
        int col = n0 + j;
        int64_t o = (int64_t)m * H + col;
        float c = tanhf(acc[j]);
        float rr = __ldg(r + o);
        if (hc) hc[o] = c;
        float hn = rr * __ldg(h + o) + (1.0f - rr) * c;
        h_out[o] = hn;
        if (h_mma) h_mma[o] = rnd ? tf32_rn(hn) : hn;
      }
    }
  }
};

}  // namespace mcrn
