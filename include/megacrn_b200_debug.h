/* Debug, probe and tuning entries of libmegacrn_b200.so: used by tests/, tools/ and bench.py's roofline section.
 * NOT part of the drop-in boundary (include/megacrn_b200.h); their behaviour may change between builds. */
#ifndef MEGACRN_B200_DEBUG_H_
#define MEGACRN_B200_DEBUG_H_

#include "megacrn_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Debug aid: mcrn_gemm on the tcgen05 engine with shared-memory stage 0 of CTA (0,0,0) dumped to dbg
 * (at least 16 K floats + 2). */
int mcrn_debug_tc_gemm(int M, int N, int K, const float* A, int lda, int trans_a,
                       const float* B, int ldb, int trans_b, float* C, int ldc, float* dbg, void* stream);

/* Round-2 probe, not on the product path (csrc/probe_mn16.cuh, tools/probe_mn16.py): one 128 x 128 x 64 tcgen05 kind::f16 tile with
 * an MN-major B operand whose shared-memory descriptor fields are given at run time.  A: device fp16 [128][64], B: device fp16
 * [64][128] (N contiguous), C: device fp32 [128][128]. */
int mcrn_debug_probe_mn16(const void* A, const void* B, float* C, unsigned lbo_bytes, unsigned sbo_bytes, unsigned layout,
                          unsigned kstep_bytes, unsigned b_major, void* stream);
/* Internal tuning knobs by name (tests / experiments): "glue_fuse" (step glue inside the gate-AGCN backward epilogue, default 0),
 * "side_chunks" (dS / dW launches per cell type, default 1), "ds_fused" (fused support-gradient kernel: 2 = fp16 operands (default), 1 = TF32, 0 = per-step GEMMs),
 * "ib_compact" (compact input block, default 1), "dw_fused" (fp16 weight-gradient kernel agcn_dw_fused_h.cuh, default 1),
 * "pdl" (programmatic dependent launch of the recurrent chain, bit mask, default 0: measured no faster under graph replay). */
int mcrn_set_option(const char* name, int value);
/* Debug: bit i forces GEMM call-site class i onto the SIMT engine (see model.cu). */
int mcrn_set_debug_mask(int mask);
/* Forward AGCN as ONE fused kernel per AGCN call (graph convolution + weight contraction + gate/update tail) where the
 * hidden width is 64 or 128: fused = 2 (default) fp16 operands / fp32 accumulate (csrc/agcn_fused_h.cuh), 1 = TF32
 * operands (csrc/agcn_fused.cuh), 0 = per-stage GEMM kernels.
 * weight_parts: 1 = the fp16 hi part of the weights (default), 2 = hi + lo residual. */
int mcrn_set_fused(int fused, int weight_parts);
/* Per-kernel timing for bench.py's roofline: while enabled, every EAGER launch of a fused AGCN kernel is bracketed by CUDA
 * events on its launching stream (launches under stream capture are not).  kernel_class = direction*8 + (HS==128 ? 4 : 0) +
 * variant; direction 0 = forward (variant 0 gate, 1 update), 1 = backward (variant 0 update-AGCN, 1 gate-AGCN).
 * mcrn_kernel_timing(1) also resets the record; _read synchronises the recorded events and returns their summed duration. */
int mcrn_kernel_timing(int enable);
int mcrn_kernel_timing_read(int kernel_class, float* ms_total, int* launches);
/* Backward data path of every AGCN as one fused kernel where the hidden width is 64 or 128: 2 (default) = fp16 operands
 * with a per-backward power-of-two loss scale (csrc/agcn_bwd_fused_h.cuh), 1 = TF32 operands (csrc/agcn_bwd_fused.cuh),
 * 0 = per-stage GEMM kernels.  Set it before the forward whose backward it governs. */
int mcrn_set_bwd_fused(int fused);
/* Debug aid (tools/fused_timeline.py): while device_slots != NULL the `which`-th fused AGCN launch after this call
 * (-1 = every launch) records clock64 timestamps of CTA (0,0) into it (512 x int64; slot map in csrc/agcn_fused.cuh).
 * NULL switches the recording off. */
int mcrn_debug_fused_timeline(long long* device_slots, int which);

/* Debug aid (tools/launch_spans.py): while device_slots != NULL every fused AGCN launch records {earliest CTA start, latest CTA
 * end} (%globaltimer, ns) into device_slots[2 i], [2 i + 1] for the i-th launch after this call (i < max_launches); the caller
 * initialises the slots to {UINT64_MAX, 0}. */
int mcrn_debug_launch_spans(unsigned long long* device_slots, int max_launches);

#ifdef __cplusplus
}
#endif
#endif /* MEGACRN_B200_DEBUG_H_ */
