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

__device__ __forceinline__ bool is16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}

// ---- epilogues ------------------------------------------------------------------
// (V == 4 fragments take a 128-bit fast path when the addresses are 16-byte aligned.)
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
    if constexpr (V == 4) {
      if (nv == 4 && is16(C + off) && (!add1 || (add1 != C && is16(add1 + off))) && (!add2 || (add2 != C && is16(add2 + off)))) {
        float v0 = alpha * acc[0], v1 = alpha * acc[1], v2 = alpha * acc[2], v3 = alpha * acc[3];
        if (add1) { float4 t = ldg4(add1 + off); v0 += t.x; v1 += t.y; v2 += t.z; v3 += t.w; }
        if (add2) { float4 t = ldg4(add2 + off); v0 += t.x; v1 += t.y; v2 += t.z; v3 += t.w; }
        if (rnd) { v0 = tf32_rn(v0); v1 = tf32_rn(v1); v2 = tf32_rn(v2); v3 = tf32_rn(v3); }
        st4(C + off, v0, v1, v2, v3);
        return;
      }
    }
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
    if constexpr (V == 4) {
      if (nv == 4 && is16(C + off)) {                       // red.global.add.v4.f32 (sm_90+)
        atomicAdd(reinterpret_cast<float4*>(C + off), make_float4(acc[0], acc[1], acc[2], acc[3]));
        return;
      }
    }
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
    if constexpr (V == 4) {
      if (nv == 4 && (W & 3) == 0) {
        const int blk = n0 / W, c = n0 - blk * W;
        float* dst = (blk == last && last_out) ? last_out + (int64_t)m * W + c
                                               : C + (int64_t)blk * blk_stride + (int64_t)m * W + c;
        if (is16(dst)) {
          const bool r_ = rnd && blk > 0 && blk != last;
          st4(dst, r_ ? tf32_rn(acc[0]) : acc[0], r_ ? tf32_rn(acc[1]) : acc[1], r_ ? tf32_rn(acc[2]) : acc[2],
              r_ ? tf32_rn(acc[3]) : acc[3]);
          return;
        }
      }
    }
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
    if constexpr (V == 4) {
      if (nv == 4 && (H & 3) == 0 && is16(h) && is16(r) && is16(zh) && (!z || is16(z))) {
        const float s0 = sigmoid_f(acc[0]), s1 = sigmoid_f(acc[1]), s2 = sigmoid_f(acc[2]), s3 = sigmoid_f(acc[3]);
        if (n0 < H) {
          const int64_t o = (int64_t)m * H + n0;
          const float4 hv = ldg4(h + o);
          if (z) st4(z + o, s0, s1, s2, s3);
          float t0 = s0 * hv.x, t1 = s1 * hv.y, t2 = s2 * hv.z, t3 = s3 * hv.w;
          if (rnd) { t0 = tf32_rn(t0); t1 = tf32_rn(t1); t2 = tf32_rn(t2); t3 = tf32_rn(t3); }
          st4(zh + o, t0, t1, t2, t3);
        } else {
          st4(r + (int64_t)m * H + (n0 - H), s0, s1, s2, s3);
        }
        return;
      }
    }
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
    if constexpr (V == 4) {
      if (nv == 4 && (H & 3) == 0 && is16(h) && is16(r) && is16(h_out) && (!hc || is16(hc)) && (!h_mma || is16(h_mma))) {
        const int64_t o = (int64_t)m * H + n0;
        const float4 hv = ldg4(h + o), rv = ldg4(r + o);
        const float c0 = tanhf(acc[0]), c1 = tanhf(acc[1]), c2 = tanhf(acc[2]), c3 = tanhf(acc[3]);
        if (hc) st4(hc + o, c0, c1, c2, c3);
        const float n0_ = rv.x * hv.x + (1.0f - rv.x) * c0, n1_ = rv.y * hv.y + (1.0f - rv.y) * c1;
        const float n2_ = rv.z * hv.z + (1.0f - rv.z) * c2, n3_ = rv.w * hv.w + (1.0f - rv.w) * c3;
        st4(h_out + o, n0_, n1_, n2_, n3_);
        if (h_mma) {
          if (rnd) st4(h_mma + o, tf32_rn(n0_), tf32_rn(n1_), tf32_rn(n2_), tf32_rn(n3_));
          else st4(h_mma + o, n0_, n1_, n2_, n3_);
        }
        return;
      }
    }
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
