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
  int use_map = 0;              // 1 = K-segment s reads A segment a_map[s] and B segment b_map[s]
  uint8_t a_map[16] = {0}, b_map[16] = {0};
  // b_sub = 2: every A segment multiplies TWO B segments (s and s + b_sub_seg: the TF32 hi and lo weight parts).
  // The tensor-core engine stages A once for both (gemm_tc_hilo); the SIMT engine walks 2*nseg segments.
  int b_sub = 1, b_sub_seg = 0;
  __host__ __device__ int nseg_a() const { return a_nseg ? a_nseg : nseg; }
  __host__ __device__ int nseg_b() const { return b_nseg ? b_nseg : nseg; }
  __host__ __device__ int total_segs() const { return nseg * b_sub; }
  __host__ __device__ int seg_a(int s) const { s = s % nseg; return use_map ? a_map[s] : s % nseg_a(); }
  __host__ __device__ int seg_b(int s) const {
    const int sub = s / nseg; s = s % nseg;
    return (use_map ? b_map[s] : s % nseg_b()) + sub * b_sub_seg;
  }
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
  // two-phase protocol of the tensor-core epilogue: all loads of a batch of rows are issued before any store
  static constexpr int NP = 2;
  __device__ __forceinline__ bool fast4(int, int, int n0) const {
    return ((ldc | c_batch | n0) & 3) == 0 && is16(C) && (!add1 || (add1 != C && is16(add1))) && (!add2 || (add2 != C && is16(add2)));
  }
  __device__ __forceinline__ void load4(int bz, int m, int n0, float4 (&p)[NP]) const {
    const int64_t off = (int64_t)bz * c_batch + (int64_t)m * ldc + n0;
    p[0] = add1 ? ldg4(add1 + off) : make_float4(0.f, 0.f, 0.f, 0.f);
    p[1] = add2 ? ldg4(add2 + off) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ void fin4(int bz, int m, int n0, const float4 (&p)[NP], const float (&acc)[4]) const {
    const int64_t off = (int64_t)bz * c_batch + (int64_t)m * ldc + n0;
    float v0 = alpha * acc[0] + p[0].x + p[1].x, v1 = alpha * acc[1] + p[0].y + p[1].y;
    float v2 = alpha * acc[2] + p[0].z + p[1].z, v3 = alpha * acc[3] + p[0].w + p[1].w;
    if (rnd) { v0 = tf32_rn(v0); v1 = tf32_rn(v1); v2 = tf32_rn(v2); v3 = tf32_rn(v3); }
    st4(C + off, v0, v1, v2, v3);
  }
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
  static constexpr int NP = 0;
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
  static constexpr int NP = 0;
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
  static constexpr int NP = 0;
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

// Column n = blk*W + c is stored at C[blk][m][c]  (the dXP block buffers); block `last` (the input block) may be
// redirected to last_out [M][W].
struct EpiBlocks {
  static constexpr int NP = 0;
  float* C; int W; int64_t blk_stride;
  int rnd = 0;                               // 1: blocks 1..last-1 (tensor-core operands downstream) are TF32-rounded
  int last = -1;
  float* last_out = nullptr;
  int blk0 = 0;                              // index of the first stored block (output column 0 belongs to block blk0)
  template <int V>
  __device__ __forceinline__ void apply(int, int m, int n0, int nv, const float (&acc)[V]) const {
    if constexpr (V == 4) {
      if (nv == 4 && (W & 3) == 0) {
        const int blk = n0 / W + blk0, c = n0 - (blk - blk0) * W;
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
        int n = n0 + j, blk = n / W + blk0, c = n - (blk - blk0) * W;
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

// Backward of the update-side propagation fused with the gate backward (tests/kernel_spec.py:cell_bwd):
//   dZH = acc + dXP0 ; dG[:, :H] = dZH*h*z(1-z) ; dG[:, H:] = dH'*(h-hc)*r(1-r) ; dh_part = dH'*r + dZH*z.
// The GEMM output is [N nodes][B*H] == flat [R][H]; dG is [R][2H].
struct EpiDG {
  const float *dxp0, *dH, *h, *z, *r, *hc;
  float *dG, *dh_part;
  int H; int64_t ld; int rnd;
  static constexpr int NP = 6;
  __device__ __forceinline__ bool fast4(int, int, int n0) const {
    return ((H | n0) & 3) == 0 && (ld & 3) == 0 && is16(dxp0) && is16(dH) && is16(h) && is16(z) && is16(r) && is16(hc) &&
           is16(dG) && is16(dh_part);
  }
  __device__ __forceinline__ void load4(int, int m, int n0, float4 (&p)[NP]) const {
    const int64_t f = (int64_t)m * ld + n0;
    p[0] = ldg4(dxp0 + f); p[1] = ldg4(z + f); p[2] = ldg4(r + f); p[3] = ldg4(h + f); p[4] = ldg4(dH + f); p[5] = ldg4(hc + f);
  }
  __device__ __forceinline__ void fin4(int, int m, int n0, const float4 (&p)[NP], const float (&acc)[4]) const {
    const int64_t flat0 = (int64_t)m * ld + n0;
    const float4 a = p[0], zz = p[1], rr = p[2], hh = p[3], dh = p[4], cc = p[5];
    const float d0 = acc[0] + a.x, d1 = acc[1] + a.y, d2 = acc[2] + a.z, d3 = acc[3] + a.w;
    float g0 = d0 * hh.x * zz.x * (1.0f - zz.x), g1 = d1 * hh.y * zz.y * (1.0f - zz.y);
    float g2 = d2 * hh.z * zz.z * (1.0f - zz.z), g3 = d3 * hh.w * zz.w * (1.0f - zz.w);
    float q0 = dh.x * (hh.x - cc.x) * rr.x * (1.0f - rr.x), q1 = dh.y * (hh.y - cc.y) * rr.y * (1.0f - rr.y);
    float q2 = dh.z * (hh.z - cc.z) * rr.z * (1.0f - rr.z), q3 = dh.w * (hh.w - cc.w) * rr.w * (1.0f - rr.w);
    if (rnd) {
      g0 = tf32_rn(g0); g1 = tf32_rn(g1); g2 = tf32_rn(g2); g3 = tf32_rn(g3);
      q0 = tf32_rn(q0); q1 = tf32_rn(q1); q2 = tf32_rn(q2); q3 = tf32_rn(q3);
    }
    const int64_t row = flat0 / H;
    const int c = (int)(flat0 - row * H);
    st4(dG + row * 2 * H + c, g0, g1, g2, g3);
    st4(dG + row * 2 * H + H + c, q0, q1, q2, q3);
    st4(dh_part + flat0, dh.x * rr.x + d0 * zz.x, dh.y * rr.y + d1 * zz.y, dh.z * rr.z + d2 * zz.z, dh.w * rr.w + d3 * zz.w);
  }
  __device__ __forceinline__ void one(int64_t flat, float acc, float& gz, float& gr, float& hp) const {
    const float dzh = acc + __ldg(dxp0 + flat), zz = __ldg(z + flat), rr = __ldg(r + flat), hh = __ldg(h + flat);
    const float dh = __ldg(dH + flat);
    gz = dzh * hh * zz * (1.0f - zz);
    gr = dh * (hh - __ldg(hc + flat)) * rr * (1.0f - rr);
    hp = dh * rr + dzh * zz;
    if (rnd) { gz = tf32_rn(gz); gr = tf32_rn(gr); }
  }
  template <int V>
  __device__ __forceinline__ void apply(int, int m, int n0, int nv, const float (&acc)[V]) const {
    const int64_t flat0 = (int64_t)m * ld + n0;
    if constexpr (V == 4) {
      if (nv == 4 && (H & 3) == 0 && (ld & 3) == 0 && is16(dxp0) && is16(dH) && is16(h) && is16(z) && is16(r) &&
          is16(hc) && is16(dG) && is16(dh_part)) {
        const float4 a = ldg4(dxp0 + flat0), zz = ldg4(z + flat0), rr = ldg4(r + flat0), hh = ldg4(h + flat0);
        const float4 dh = ldg4(dH + flat0), cc = ldg4(hc + flat0);
        const float d0 = acc[0] + a.x, d1 = acc[1] + a.y, d2 = acc[2] + a.z, d3 = acc[3] + a.w;
        float g0 = d0 * hh.x * zz.x * (1.0f - zz.x), g1 = d1 * hh.y * zz.y * (1.0f - zz.y);
        float g2 = d2 * hh.z * zz.z * (1.0f - zz.z), g3 = d3 * hh.w * zz.w * (1.0f - zz.w);
        float q0 = dh.x * (hh.x - cc.x) * rr.x * (1.0f - rr.x), q1 = dh.y * (hh.y - cc.y) * rr.y * (1.0f - rr.y);
        float q2 = dh.z * (hh.z - cc.z) * rr.z * (1.0f - rr.z), q3 = dh.w * (hh.w - cc.w) * rr.w * (1.0f - rr.w);
        if (rnd) {
          g0 = tf32_rn(g0); g1 = tf32_rn(g1); g2 = tf32_rn(g2); g3 = tf32_rn(g3);
          q0 = tf32_rn(q0); q1 = tf32_rn(q1); q2 = tf32_rn(q2); q3 = tf32_rn(q3);
        }
        const int64_t row = flat0 / H;
        const int c = (int)(flat0 - row * H);
        st4(dG + row * 2 * H + c, g0, g1, g2, g3);
        st4(dG + row * 2 * H + H + c, q0, q1, q2, q3);
        st4(dh_part + flat0, dh.x * rr.x + d0 * zz.x, dh.y * rr.y + d1 * zz.y, dh.z * rr.z + d2 * zz.z, dh.w * rr.w + d3 * zz.w);
        return;
      }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (j < nv) {
        const int64_t flat = flat0 + j, row = flat / H;
        const int c = (int)(flat - row * H);
        float gz, gr, hp;
        one(flat, acc[j], gz, gr, hp);
        dG[row * 2 * H + c] = gz;
        dG[row * 2 * H + H + c] = gr;
        dh_part[flat] = hp;
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
  static constexpr int NP = 1;
  __device__ __forceinline__ bool fast4(int, int, int n0) const {
    return ((H | n0) & 3) == 0 && is16(h) && is16(r) && is16(zh) && (!z || is16(z));
  }
  __device__ __forceinline__ void load4(int, int m, int n0, float4 (&p)[NP]) const {
    p[0] = (n0 < H) ? ldg4(h + (int64_t)m * H + n0) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __device__ __forceinline__ void fin4(int, int m, int n0, const float4 (&p)[NP], const float (&acc)[4]) const {
    const float s0 = sigmoid_f(acc[0]), s1 = sigmoid_f(acc[1]), s2 = sigmoid_f(acc[2]), s3 = sigmoid_f(acc[3]);
    if (n0 < H) {
      const int64_t o = (int64_t)m * H + n0;
      if (z) st4(z + o, s0, s1, s2, s3);
      float t0 = s0 * p[0].x, t1 = s1 * p[0].y, t2 = s2 * p[0].z, t3 = s3 * p[0].w;
      if (rnd) { t0 = tf32_rn(t0); t1 = tf32_rn(t1); t2 = tf32_rn(t2); t3 = tf32_rn(t3); }
      st4(zh + o, t0, t1, t2, t3);
    } else {
      st4(r + (int64_t)m * H + (n0 - H), s0, s1, s2, s3);
    }
  }
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
  static constexpr int NP = 2;
  __device__ __forceinline__ bool fast4(int, int, int n0) const {
    return ((H | n0) & 3) == 0 && is16(h) && is16(r) && is16(h_out) && (!hc || is16(hc)) && (!h_mma || is16(h_mma));
  }
  __device__ __forceinline__ void load4(int, int m, int n0, float4 (&p)[NP]) const {
    const int64_t o = (int64_t)m * H + n0;
    p[0] = ldg4(h + o);
    p[1] = ldg4(r + o);
  }
  __device__ __forceinline__ void fin4(int, int m, int n0, const float4 (&p)[NP], const float (&acc)[4]) const {
    const int64_t o = (int64_t)m * H + n0;
    const float4 hv = p[0], rv = p[1];
    const float c0 = tanhf(acc[0]), c1 = tanhf(acc[1]), c2 = tanhf(acc[2]), c3 = tanhf(acc[3]);
    if (hc) st4(hc + o, c0, c1, c2, c3);
    const float n0_ = rv.x * hv.x + (1.0f - rv.x) * c0, n1_ = rv.y * hv.y + (1.0f - rv.y) * c1;
    const float n2_ = rv.z * hv.z + (1.0f - rv.z) * c2, n3_ = rv.w * hv.w + (1.0f - rv.w) * c3;
    st4(h_out + o, n0_, n1_, n2_, n3_);
    if (h_mma) {
      if (rnd) st4(h_mma + o, tf32_rn(n0_), tf32_rn(n1_), tf32_rn(n2_), tf32_rn(n3_));
      else st4(h_mma + o, n0_, n1_, n2_, n3_);
    }
  }
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
