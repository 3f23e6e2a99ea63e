/*
 * megacrn_b200.h -- C ABI of libmegacrn_b200.so (sm_100a only).
 *
 * The reference (deepkashiwa20/MegaCRN) has no FFI of its own: its only boundary
 * for this path is the Python class `MegaCRN(nn.Module)` (model/MegaCRN.py:116-194).
 * Each entry point below names the reference code it replaces; the Python module
 * megacrn_b200/MegaCRN.py binds them with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous row-major fp32 unless marked
 *     "host"; buffers are owned by the caller (PyTorch allocations); the library
 *     never allocates device memory, never frees, never retains a pointer.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call
 *     synchronises the device except the *_host convenience entries.
 *   - return value: 0 on success, a negative mcrn_status otherwise; the message is
 *     available from mcrn_last_error() (thread-local).  Nothing throws or exits.
 *   - there is no CPU fallback: on a machine without an sm_100 device the compute
 *     entries return MCRN_ERR_NO_DEVICE.
 */
#ifndef MEGACRN_B200_H_
#define MEGACRN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCRN_ABI_VERSION 1

typedef enum mcrn_status {
  MCRN_OK = 0,
  MCRN_ERR_BAD_DIMS = -1,      /* unsupported / inconsistent dimensions            */
  MCRN_ERR_BAD_POINTER = -2,   /* null or misaligned (16 B) pointer                */
  MCRN_ERR_WORKSPACE = -3,     /* workspace smaller than mcrn_workspace_bytes()    */
  MCRN_ERR_NO_DEVICE = -4,     /* no CUDA device of compute capability 10.x        */
  MCRN_ERR_CUDA = -5,          /* a CUDA runtime / driver call failed              */
  MCRN_ERR_STATE = -6          /* backward called on a workspace without a forward */
} mcrn_status;

/* Constructor arguments of the reference model (model/MegaCRN.py:117-118) plus the
 * per-call batch geometry (x.shape, model/MegaCRN.py:168). */
typedef struct mcrn_dims {
  int32_t batch;        /* B      : x.shape[0]                                       */
  int32_t num_nodes;    /* N                                                         */
  int32_t seq_len;      /* T_in   : x.shape[1]                                       */
  int32_t horizon;      /* T_out                                                     */
  int32_t input_dim;    /* Cin    : x.shape[3]                                       */
  int32_t output_dim;   /* Cout                                                      */
  int32_t ycov_dim;
  int32_t rnn_units;    /* H                                                         */
  int32_t num_layers;   /* L, 1..MCRN_MAX_LAYERS; L > 1 goes through mcrn_forward_layers / mcrn_backward_layers */
  int32_t cheb_k;       /* >= 2                                                      */
  int32_t mem_num;      /* M                                                         */
  int32_t mem_dim;      /* d                                                         */
} mcrn_dims;

/* The 14 parameter tensors of the reference state_dict, reference layout
 * (shapes from model/MegaCRN.py:11-12, :151-154, :144).
 * The same struct carries gradients for mcrn_backward. */
typedef struct mcrn_params {
  float* memory;        /* memory.Memory                       [M, d]                */
  float* wq;            /* memory.Wq                           [H, d]                */
  float* we1;           /* memory.We1                          [N, M]                */
  float* we2;           /* memory.We2                          [N, M]                */
  float* enc_gate_w;    /* encoder.dcrnn_cells.0.gate.weights  [2k(Cin+H), 2H]       */
  float* enc_gate_b;    /*                     ...gate.bias    [2H]                  */
  float* enc_update_w;  /*                     ...update.weights [2k(Cin+H), H]      */
  float* enc_update_b;  /*                     ...update.bias  [H]                   */
  float* dec_gate_w;    /* decoder.dcrnn_cells.0.gate.weights  [2k(Cout+ycov+D), 2D] */
  float* dec_gate_b;    /*                                     [2D],  D = H + d      */
  float* dec_update_w;  /*                                     [2k(Cout+ycov+D), D]  */
  float* dec_update_b;  /*                                     [D]                   */
  float* proj_w;        /* proj.0.weight                       [Cout, D]             */
  float* proj_b;        /* proj.0.bias                         [Cout]                */
} mcrn_params;

/* Stacked cells (num_layers > 1; model/MegaCRN.py:62-63, :71-78, :100-101, :109-112): the parameters of
 * encoder.dcrnn_cells.{i} / decoder.dcrnn_cells.{i} for ONE layer i >= 1.  Their input is the state of the layer below, so
 * dim_in = H for the encoder cell and D = H + d for the decoder cell.  The same struct carries their gradients. */
#define MCRN_MAX_LAYERS 4
typedef struct mcrn_layer_params {
  float* enc_gate_w;    /* encoder.dcrnn_cells.{i}.gate.weights    [2k(H+H), 2H]   */
  float* enc_gate_b;    /*                      ...gate.bias      [2H]            */
  float* enc_update_w;  /*                      ...update.weights [2k(H+H), H]    */
  float* enc_update_b;  /*                      ...update.bias    [H]             */
  float* dec_gate_w;    /* decoder.dcrnn_cells.{i}.gate.weights    [2k(D+D), 2D]   */
  float* dec_gate_b;    /*                                        [2D]            */
  float* dec_update_w;  /*                                        [2k(D+D), D]    */
  float* dec_update_b;  /*                                        [D]             */
} mcrn_layer_params;

/* Flags for mcrn_forward. */
#define MCRN_FWD_SAVE_FOR_BACKWARD 1u   /* keep per-step activations in the workspace */
#define MCRN_FWD_REUSE_PROLOGUE    2u   /* eval fast path: the supports (model/MegaCRN.py:169-173, :19-23), the folded
                                         * weights and their operand copies depend on the parameters only; the caller asserts that the
                                         * workspace still holds them from a previous mcrn_forward with the SAME dims, flags, parameter
                                         * values and library mode (mcrn_mode_epoch unchanged), and the prologue kernels are skipped */

int mcrn_abi_version(void);
const char* mcrn_last_error(void);

/* 0 if a usable sm_100 device is current, MCRN_ERR_NO_DEVICE otherwise. */
int mcrn_device_ok(void);

/* Bytes of caller-provided device workspace needed by mcrn_forward (+ mcrn_backward
 * when `flags & MCRN_FWD_SAVE_FOR_BACKWARD`).  Returns 0 and sets the error on bad dims. */
size_t mcrn_workspace_bytes(const mcrn_dims* dims, uint32_t flags);

/* Whole forward of the reference model: MegaCRN.forward (model/MegaCRN.py:168-194)
 * = supports prologue (:169-173) + ADCRNN_Encoder.forward (:65-83) over AGCRNCell
 * (:38-48) / AGCN (:16-28) + query_memory (:159-166) + decoder loop with proj and
 * scheduled sampling (:181-192).
 *   x       [B, T_in,  N, Cin]      y_cov  [B, T_out, N, ycov]
 *   labels  [B, T_out, N, Cout] or NULL when no step is teacher-forced
 *   teacher_forcing  HOST array of T_out bytes (or NULL = all zero): byte t != 0
 *           means the coin flip of model/MegaCRN.py:189-191 at step t selected
 *           labels[:, t] as the next decoder input (the host draws the coins, in
 *           the reference's np.random order, before calling).
 *   outputs: output [B, T_out, N, Cout]; h_att, query, pos, neg [B, N, d]. */
int mcrn_forward(const mcrn_dims* dims, const mcrn_params* params,
                 const float* x, const float* y_cov, const float* labels,
                 const uint8_t* teacher_forcing,
                 float* output, float* h_att, float* query, float* pos, float* neg,
                 void* workspace, size_t workspace_bytes, uint32_t flags, void* stream);

/* Backward of the same (what autograd does for the reference at
 * model/traintest_MegaCRN.py:128): given upstream gradients of the five outputs
 * (any may be NULL = zero), writes (not accumulates) the 14 parameter gradients in
 * reference layout into `grads`.  Must follow an mcrn_forward on the same workspace
 * with MCRN_FWD_SAVE_FOR_BACKWARD, same dims/params/inputs/teacher_forcing. */
int mcrn_backward(const mcrn_dims* dims, const mcrn_params* params,
                  const float* x, const float* y_cov, const float* labels,
                  const uint8_t* teacher_forcing,
                  const float* d_output, const float* d_h_att, const float* d_query,
                  const float* d_pos, const float* d_neg,
                  const mcrn_params* grads,
                  void* workspace, size_t workspace_bytes, void* stream);

/* num_layers >= 1: mcrn_forward / mcrn_backward with the stacked cells of layers 1 .. num_layers-1 in `upper`
 * (an array of num_layers-1 structs; NULL when num_layers == 1, in which case these ARE mcrn_forward / mcrn_backward).
 * ADCRNN_Encoder.forward runs layer by layer over the whole sequence (model/MegaCRN.py:71-78), ADCRNN_Decoder.forward
 * runs the stack once per step (:109-112), every decoder layer starts from the same [h_T | h_att] (:181) and the
 * projection reads the top layer (:186).  Layers >= 1 run on the per-stage GEMM engine (their AGCN operand is the
 * 2*width concatenation [state below | own state]); with num_layers > 1 layer 0 does too.  MCRN_FWD_REUSE_PROLOGUE is
 * ignored for num_layers > 1.  mcrn_forward / mcrn_backward themselves reject num_layers > 1. */
int mcrn_forward_layers(const mcrn_dims* dims, const mcrn_params* params, const mcrn_layer_params* upper,
                        const float* x, const float* y_cov, const float* labels,
                        const uint8_t* teacher_forcing,
                        float* output, float* h_att, float* query, float* pos, float* neg,
                        void* workspace, size_t workspace_bytes, uint32_t flags, void* stream);
int mcrn_backward_layers(const mcrn_dims* dims, const mcrn_params* params, const mcrn_layer_params* upper,
                         const float* x, const float* y_cov, const float* labels,
                         const uint8_t* teacher_forcing,
                         const float* d_output, const float* d_h_att, const float* d_query,
                         const float* d_pos, const float* d_neg,
                         const mcrn_params* grads, const mcrn_layer_params* upper_grads,
                         void* workspace, size_t workspace_bytes, void* stream);

/* The caller-side loss of the training step, fused:
 * loss = masked_mae(output*std+mean, labels*std+mean)            (model/utils.py:126-133)
 *      + lamb  * TripletMarginLoss(margin=1)(query, pos, neg)    (model/traintest_MegaCRN.py:121-123)
 *      + lamb1 * MSELoss(query, pos)                             (:122-124)
 * with pos/neg treated as constants (the trainer detaches them).  Writes the scalar
 * loss to loss_out[0] and d(loss)/d(output), d(loss)/d(query) (either may be NULL).
 * workspace: at least 256 bytes of device scratch. */
int mcrn_trainer_loss(const mcrn_dims* dims, const float* output, const float* labels,
                      const float* query, const float* pos, const float* neg,
                      float scaler_mean, float scaler_std, float lamb, float lamb1,
                      float* loss_out, float* d_output, float* d_query,
                      void* workspace, size_t workspace_bytes, void* stream);

/* Data-parallel form of the loss.  masked_mae's normaliser `mask.mean()`
 * (model/utils.py:127-128) is a property of the GLOBAL batch: with the batch sharded over W ranks, the W-rank
 * average of the per-rank losses equals the single-process loss only if every rank divides by
 * (global count of labels != 0) / W instead of its local count.  mcrn_mask_count adds this rank's count of
 * (labels*std+mean != 0) to count_out[0] (device; the caller zeroes it, all-reduces it and divides by W);
 * mcrn_trainer_loss_dp is mcrn_trainer_loss with that device scalar as the normaliser (NULL: local count). */
int mcrn_mask_count(const float* labels, int64_t n, float scaler_mean, float scaler_std, float* count_out, void* stream);
int mcrn_trainer_loss_dp(const mcrn_dims* dims, const float* output, const float* labels,
                         const float* query, const float* pos, const float* neg,
                         float scaler_mean, float scaler_std, float lamb, float lamb1,
                         const float* mask_count,
                         float* loss_out, float* d_output, float* d_query,
                         void* workspace, size_t workspace_bytes, void* stream);

/* The optimiser half of the training step, fused over the 14 parameter tensors:
 * torch.nn.utils.clip_grad_norm_(parameters, max_grad_norm)   (model/traintest_MegaCRN.py:129; max_grad_norm <= 0: no clipping)
 * followed by torch.optim.Adam(lr, betas=(beta1, beta2), eps).step()   (:104, :130; no weight decay, no amsgrad).
 * params / grads / exp_avg / exp_avg_sq: device tensors in mcrn_params layout (shapes from dims).
 * dev_state: 4 device floats = { step count (incremented by the call), learning rate (read), total gradient norm (written),
 * clip coefficient (written) } -- device-resident so the call is CUDA-graph capturable and needs no host sync. */
int mcrn_adam_step(const mcrn_dims* dims, const mcrn_params* params, const mcrn_params* grads,
                   const mcrn_params* exp_avg, const mcrn_params* exp_avg_sq, float* dev_state,
                   float beta1, float beta2, float eps, float max_grad_norm, void* stream);

/* Same over the 14 + 8*(num_layers-1) tensors of a model with stacked cells: ONE gradient norm over all of them (what
 * clip_grad_norm_(model.parameters()) computes) and one update.  The upper_* arrays hold num_layers-1 structs (NULL when
 * num_layers == 1, in which case this IS mcrn_adam_step). */
int mcrn_adam_step_layers(const mcrn_dims* dims, const mcrn_params* params, const mcrn_layer_params* upper,
                          const mcrn_params* grads, const mcrn_layer_params* upper_grads,
                          const mcrn_params* exp_avg, const mcrn_layer_params* upper_exp_avg,
                          const mcrn_params* exp_avg_sq, const mcrn_layer_params* upper_exp_avg_sq,
                          float* dev_state, float beta1, float beta2, float eps, float max_grad_norm, void* stream);

/* HOST-buffer convenience entries (what a non-PyTorch caller would bind): every
 * pointer in params/grads/x/... is a HOST pointer; the library stages through the
 * caller-provided DEVICE workspace, which must be at least
 * mcrn_host_workspace_bytes() long, and synchronises `stream` before returning. */
size_t mcrn_host_workspace_bytes(const mcrn_dims* dims, uint32_t flags);
int mcrn_forward_host(const mcrn_dims* dims, const mcrn_params* host_params,
                      const float* x, const float* y_cov, const float* labels,
                      const uint8_t* teacher_forcing,
                      float* output, float* h_att, float* query, float* pos, float* neg,
                      void* device_workspace, size_t workspace_bytes, uint32_t flags, void* stream);

/* Stage-level entries used by the unit tests (same arithmetic the whole-model entries
 * run; see tests/test_gpu_parity.py: test_gemm_engine, test_supports_stage_matches_spec). */

/* Supports prologue: model/MegaCRN.py:169-173 and the Chebyshev set of :19-23 hoisted.
 * supports_out [KS, N, ld] with KS = 2*(cheb_k-1), ld = mcrn_support_ld(N). */
int mcrn_support_ld(int num_nodes);
int mcrn_supports_fwd(const mcrn_dims* dims, const float* memory, const float* we1, const float* we2,
                      float* supports_out, void* workspace, size_t workspace_bytes, void* stream);

/* Same, also returning the TF32-rounded copy the tensor-core GEMMs read (tests). */
int mcrn_supports_fwd2(const mcrn_dims* dims, const float* memory, const float* we1, const float* we2,
                       float* supports_out, float* supports_tc_out, void* workspace, size_t workspace_bytes,
                       void* stream);

/* Generic row-major GEMM  C[M,N] = A[M,K] * B[K,N]  on the library's GEMM engine
 * (the engine every stage uses).  trans_a / trans_b: the operand is stored
 * transposed ([K,M] / [N,K]).  engine: 0 = library default, 1 = SIMT fp32,
 * 2 = tcgen05 TF32.  */
int mcrn_gemm(int M, int N, int K, const float* A, int lda, int trans_a,
              const float* B, int ldb, int trans_b, float* C, int ldc,
              int engine, void* stream);

/* Number of kernels the library launched since process start (for bench.py's
 * `gpu_launches`). */
uint64_t mcrn_launch_count(void);

/* 0 = default (tcgen05 TF32 where the shape allows, SIMT otherwise), 1 = force SIMT fp32. */
int mcrn_set_engine(int engine);
int mcrn_get_engine(void);
/* Counter incremented by every mcrn_set_* call: part of the validity key of MCRN_FWD_REUSE_PROLOGUE. */
uint64_t mcrn_mode_epoch(void);

/* Debug, probe and tuning entries (mcrn_debug_*, mcrn_set_option, mcrn_set_fused, mcrn_kernel_timing, ...) are declared in
 * megacrn_b200_debug.h: tests, tools and bench.py's roofline use them; a binding of the product path does not need them. */

#ifdef __cplusplus
}
#endif
#endif /* MEGACRN_B200_H_ */
