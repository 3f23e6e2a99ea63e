// Stacked cells: num_layers > 1 (model/MegaCRN.py:62-63, :71-78, :100-101, :109-112).  Included by model.cu (uses its
// per-stage building blocks: propagate, make_dxp, propagate_T, acc_ds, acc_dw_all, cell_forward, cell_backward).
//
// A cell of layer l >= 1 takes the Hs-wide state of the layer below as its input, which the input block of the layer-0
// cells (<= Hs columns for ALL input channels of all supports) cannot carry.  Its AGCN is instead run over the concatenated
// operand V = [x_in | h] of width Kw = 2*Hs with no separate input channels: the reference's weight rows are already
// ordered (support block, [input | state]) (model/MegaCRN.py:42, :24-27), so k_fold_weights(cin = 0, hs = Kw) is the
// folded weight of that operand, the input block degenerates to the bias column, and k_unfold_grads(cin = 0, hs = Kw)
// maps the accumulated gradient back.  Algebra and BPTT order: tests/kernel_spec.py (cell_fwd_wide, cell_bwd_wide,
// model_fwd_layers, model_bwd_layers), proven against autograd of the reference formulation in fp64.
//
// With num_layers > 1 every layer runs on the per-stage GEMM engine (TF32 tcgen05 GEMMs, or SIMT fp32 with
// mcrn_set_engine(1)) on one stream; the fused fp16 kernels are the num_layers == 1 path.
#pragma once
// (included INSIDE namespace mcrn of model.cu, after backward_impl)

// every predicate that selects a fused kernel reads these globals: inside this scope all layers take the per-stage path
struct PerStageScope {
  int f, b, d;
  PerStageScope() : f(g_fused), b(g_bwd_fused), d(g_ds_fused) { g_fused = 0; g_bwd_fused = 0; g_ds_fused = 0; }
  ~PerStageScope() { g_fused = f; g_bwd_fused = b; g_ds_fused = d; }
};

struct UpperW { const float *wg, *wu; int Hs; };      // folded [hi|lo][NB+1][2Hs][2Hs] / [..][2Hs][Hs]
struct UpperBufs {
  float *xpg, *xpu;      // [NB+1][R][2Hs]: block 0 = [x_in | h] / [x_in | z*h], blocks 1..KS propagated, block NB = bias column
  float *z, *r, *hc;     // [R][Hs] (z, hc null in eval)
  const float* hx;       // exact fp32 input state [R][Hs]
};

static UpperBufs upper_enc_bufs(const Geo& g, const Plan& p, float* ws, int l, int t) {
  const UpperPlan& u = p.up[l - 1];
  const int s = t % p.up_slots_e;
  UpperBufs b;
  b.xpg = ws + u.e_xpg + p.up_e_xp_sz * s;
  b.xpu = ws + u.e_xpu + p.up_e_xp_sz * s;
  b.z = p.save ? ws + u.e_z + p.enc_v_sz * s : nullptr;
  b.r = ws + u.e_r + p.enc_v_sz * s;
  b.hc = p.save ? ws + u.e_hc + p.enc_v_sz * s : nullptr;
  b.hx = t > 0 ? ws + u.e_hseq + p.enc_v_sz * (size_t)(t - 1) : ws + p.up_zero;      // zero initial state (:50-51, :72)
  return b;
}
static UpperBufs upper_dec_bufs(const Geo& g, const Plan& p, float* ws, int l, int t) {
  const UpperPlan& u = p.up[l - 1];
  const int s = t % p.up_slots_d;
  UpperBufs b;
  b.xpg = ws + u.d_xpg + p.up_d_xp_sz * s;
  b.xpu = ws + u.d_xpu + p.up_d_xp_sz * s;
  b.z = p.save ? ws + u.d_z + p.dec_v_sz * s : nullptr;
  b.r = ws + u.d_r + p.dec_v_sz * s;
  b.hc = p.save ? ws + u.d_hc + p.dec_v_sz * s : nullptr;
  b.hx = t > 0 ? ws + u.d_hseq + p.dec_v_sz * (size_t)(t - 1) : ws + p.dec_hx;         // every layer starts from [h_T | h_att] (:181)
  return b;
}

// forward of one stacked cell (model/MegaCRN.py:38-48 with x = the state of the layer below)
static int cell_forward_wide(const Geo& g, const float* S, const UpperW& w, const UpperBufs& b, const float* x_in,
                             float* zh_tmp, float* h_out, cudaStream_t st) {
  const int Hs = w.Hs, Kw = 2 * Hs, NBX = g.NB + 1, rnd = tf32_mode();
  const int64_t nK = g.R * Kw;
  MCRN_LAUNCH(k_concat2, ew_grid(nK), 256, 0, st, x_in, b.hx, b.xpg, g.R, Hs, rnd);
  MCRN_LAUNCH(k_build_input_block, ew_grid(nK), 256, 0, st, (const float*)nullptr, (int64_t)0, (int64_t)0, g.NB, 0, g.B, g.R, Kw,
              rnd, b.xpg + (int64_t)g.NB * nK, b.xpu + (int64_t)g.NB * nK, (__half*)nullptr);
  MCRN_TRY(propagate(g, S, b.xpg, Kw, st));
  {  // gate AGCN + sigmoid + z*h                                   model/MegaCRN.py:42-45
    GemmDesc q;
    q.A = b.xpg; q.a_row = Kw; q.a_k = 1; q.a_seg = nK; q.nseg = NBX; q.Kseg = Kw; q.M = (int)g.R;
    if (rnd) hilo(q, NBX);
    q.B = w.wg; q.b_seg = (int64_t)Kw * 2 * Hs; q.b_k = 2 * Hs; q.b_n = 1; q.N = 2 * Hs; q.prec_exact = dbg_exact(1);
    EpiGate e{Hs, b.hx, b.z, b.r, zh_tmp, rnd};
    MCRN_TRY(gemm(q, e, st));
  }
  MCRN_LAUNCH(k_concat2, ew_grid(nK), 256, 0, st, x_in, (const float*)zh_tmp, b.xpu, g.R, Hs, rnd);
  MCRN_TRY(propagate(g, S, b.xpu, Kw, st));
  {  // update AGCN + tanh + blend                                  model/MegaCRN.py:46-47
    GemmDesc q;
    q.A = b.xpu; q.a_row = Kw; q.a_k = 1; q.a_seg = nK; q.nseg = NBX; q.Kseg = Kw; q.M = (int)g.R;
    if (rnd) hilo(q, NBX);
    q.B = w.wu; q.b_seg = (int64_t)Kw * Hs; q.b_k = Hs; q.b_n = 1; q.N = Hs; q.prec_exact = dbg_exact(1);
    EpiUpdate e{Hs, b.hx, b.r, b.hc, h_out, nullptr, rnd};
    MCRN_TRY(gemm(q, e, st));
  }
  return MCRN_OK;
}

// backward of one stacked cell (tests/kernel_spec.py: cell_bwd_wide).  dH: in = grad of h', out = grad of h (in place);
// dx = grad of x_in (written); dU / dG: this step's slices (kept for the weight gradients); dS accumulates.
static int cell_backward_wide(const Geo& g, const Plan& p, float* ws, const float* S, const UpperW& w, const UpperBufs& b,
                              float* dU, float* dG, float* dH, float* dx, cudaStream_t st) {
  const int Hs = w.Hs, Kw = 2 * Hs, rnd = tf32_mode();
  const int64_t nH = g.R * Hs, nK = g.R * Kw;
  float *dXP = ws + p.up_dXP, *dXP2 = ws + p.up_dXP2, *dV0 = ws + p.up_dV0, *dhp = ws + p.up_dhp, *dxa = ws + p.up_dxa;
  float* dS = ws + p.dS;
  MCRN_LAUNCH(k_bwd_du, ew_grid(nH), 256, 0, st, dH, b.r, b.hc, dU, nH, rnd);
  // update AGCN: dXP = dU Wu^T ; d[x_in | z*h] = dXP_0 + sum_k S_k^T dXP_k
  MCRN_TRY(make_dxp(g, dU, Hs, w.wu, Kw, dXP, nullptr, st));
  MCRN_TRY(acc_ds(g, dXP + nK, (int64_t)g.B * Kw, b.xpu, (int64_t)g.B * Kw, g.B * Kw, dS, st));
  MCRN_TRY(propagate_T(g, S, dXP, Kw, nullptr, dV0, st));
  MCRN_LAUNCH(k_wide_gate_bwd, ew_grid(nH), 256, 0, st, dV0, dH, b.hx, b.z, b.r, b.hc, dG, dhp, dxa, g.R, Hs, rnd);
  // gate AGCN: dXP2 = dG Wg^T ; d[x_in | h] = dXP2_0 + sum_k S_k^T dXP2_k
  MCRN_TRY(make_dxp(g, dG, 2 * Hs, w.wg, Kw, dXP2, nullptr, st));
  MCRN_TRY(acc_ds(g, dXP2 + nK, (int64_t)g.B * Kw, b.xpg, (int64_t)g.B * Kw, g.B * Kw, dS, st));
  MCRN_TRY(propagate_T(g, S, dXP2, Kw, nullptr, dV0, st));
  MCRN_LAUNCH(k_wide_finish, ew_grid(nH), 256, 0, st, dV0, dhp, dxa, dH, dx, g.R, Hs);
  return MCRN_OK;
}

static int fold_upper_weights(const Geo& g, const Plan& p, const mcrn_layer_params* up, float* ws, cudaStream_t st) {
  const int sp = tf32_mode();
  for (int l = 1; l < g.L; ++l) {
    const UpperPlan& u = p.up[l - 1];
    const mcrn_layer_params& q = up[l - 1];
    MCRN_LAUNCH(k_fold_weights, 128, 256, 0, st, q.enc_gate_w, q.enc_gate_b, ws + u.e_wg, 0, 2 * g.H, 2 * g.H, g.cheb_k, sp);
    MCRN_LAUNCH(k_fold_weights, 128, 256, 0, st, q.enc_update_w, q.enc_update_b, ws + u.e_wu, 0, 2 * g.H, g.H, g.cheb_k, sp);
    MCRN_LAUNCH(k_fold_weights, 128, 256, 0, st, q.dec_gate_w, q.dec_gate_b, ws + u.d_wg, 0, 2 * g.D, 2 * g.D, g.cheb_k, sp);
    MCRN_LAUNCH(k_fold_weights, 128, 256, 0, st, q.dec_update_w, q.dec_update_b, ws + u.d_wu, 0, 2 * g.D, g.D, g.cheb_k, sp);
  }
  return MCRN_OK;
}

// output of encoder layer l at step t (l == 0: the layer-0 state buffers of forward_impl)
static const float* enc_layer_out(const Geo& g, const Plan& p, float* ws, int l, int t) {
  if (l == 0) return (t + 1 < g.T_in) ? enc_bufs(g, p, ws, t + 1).hx : ws + p.h_enc;
  return ws + p.up[l - 1].e_hseq + p.enc_v_sz * (size_t)t;
}
static const float* dec_layer_out(const Geo& g, const Plan& p, float* ws, int l, int t) {
  if (l == 0) return (t + 1 < g.T_out) ? dec_bufs(g, p, ws, t + 1).hx : ws + p.h_dec_last;
  return ws + p.up[l - 1].d_hseq + p.dec_v_sz * (size_t)t;
}

int forward_impl_layers(const Geo& g, const Plan& p, const mcrn_params* prm, const mcrn_layer_params* up, const float* x,
                        const float* y_cov, const float* labels, const uint8_t* tf, float* output, float* h_att, float* query,
                        float* pos, float* neg, float* ws, cudaStream_t st) {
  PerStageScope per_stage;
  float* S = ws + p.Sr;
  const int sp = tf32_mode();
  // ---- parameter-only prologue: folded weights of every layer, supports (:169-173, :19-23) ----
  MCRN_LAUNCH(k_fold_weights, 128, 256, 0, st, prm->enc_gate_w, prm->enc_gate_b, ws + p.e_wg, g.Cin, g.H, 2 * g.H, g.cheb_k, sp);
  MCRN_LAUNCH(k_fold_weights, 128, 256, 0, st, prm->enc_update_w, prm->enc_update_b, ws + p.e_wu, g.Cin, g.H, g.H, g.cheb_k, sp);
  MCRN_LAUNCH(k_fold_weights, 128, 256, 0, st, prm->dec_gate_w, prm->dec_gate_b, ws + p.d_wg, g.Cdec, g.D, 2 * g.D, g.cheb_k, sp);
  MCRN_LAUNCH(k_fold_weights, 128, 256, 0, st, prm->dec_update_w, prm->dec_update_b, ws + p.d_wu, g.Cdec, g.D, g.D, g.cheb_k, sp);
  MCRN_TRY(fold_upper_weights(g, p, up, ws, st));
  MCRN_TRY(supports_forward(g, p, ws, prm->memory, prm->we1, prm->we2, ws + p.S, ws + p.Sr, st));
  MCRN_CUDA_OK(cudaMemsetAsync(ws + p.up_zero, 0, (size_t)g.R * g.D * sizeof(float), st));
  // ---- encoder, layer by layer over the whole sequence (:71-78) ----
  {
    const int64_t n_in = (int64_t)g.N * g.T_in * g.B * g.Cin;
    MCRN_LAUNCH(k_stage_encoder_input, ew_grid(n_in), 256, 0, st, x, ws + p.enc_xpin, g.B, g.T_in, g.N, g.Cin, sp);
    MCRN_TRY(propagate_in(g, S, ws + p.enc_xpin, (int64_t)g.N * g.T_in * g.B * g.Cin, (int64_t)g.T_in * g.B * g.Cin,
                          g.T_in * g.B * g.Cin, st));
    MCRN_CUDA_OK(cudaMemsetAsync(ws + p.enc_xpg, 0, (size_t)g.R * g.H * sizeof(float), st));
    MCRN_CUDA_OK(cudaMemsetAsync(ws + p.enc_hx, 0, (size_t)g.R * g.H * sizeof(float), st));
    CellW w = enc_w(g, p, ws);
    for (int t = 0; t < g.T_in; ++t) {
      CellBufs b = enc_bufs(g, p, ws, t);
      const bool last = (t + 1 == g.T_in);
      float* h_out = last ? ws + p.h_enc : enc_bufs(g, p, ws, t + 1).hx;
      float* h_mma = last ? nullptr : enc_bufs(g, p, ws, t + 1).xpg;
      MCRN_TRY(cell_forward(g, S, w, b, h_out, h_mma, st, nullptr));
    }
    for (int l = 1; l < g.L; ++l) {
      const UpperW uw{ws + p.up[l - 1].e_wg, ws + p.up[l - 1].e_wu, g.H};
      for (int t = 0; t < g.T_in; ++t) {
        UpperBufs b = upper_enc_bufs(g, p, ws, l, t);
        MCRN_TRY(cell_forward_wide(g, S, uw, b, enc_layer_out(g, p, ws, l - 1, t), ws + p.up_zh,
                                   ws + p.up[l - 1].e_hseq + p.enc_v_sz * (size_t)t, st));
      }
    }
  }
  // ---- memory query on the top layer's last state (:176-179) + the decoder's initial state ----
  const float* h_top = enc_layer_out(g, p, ws, g.L - 1, g.T_in - 1);
  {
    CellBufs b0 = dec_bufs(g, p, ws, 0);
    const size_t shm = (8 * (size_t)(g.H + g.d + g.M) + (size_t)g.M * (g.d + 1)) * sizeof(float);
    static bool mq_attr = false;
    if (!mq_attr) {
      MCRN_CUDA_OK(cudaFuncSetAttribute(k_memory_query, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      mq_attr = true;
    }
    if (shm > 160 * 1024) { set_error("memory query: rnn_units * mem_dim too large for the shared-memory staging (%zu bytes)", shm); return MCRN_ERR_BAD_DIMS; }
    MCRN_LAUNCH(k_memory_query, (int)ceil_div64(g.R, 8 * MQ_ROWS_PER_WARP), 256, shm, st, h_top, prm->wq, prm->memory,
                ws + p.mq_q, ws + p.mq_att, reinterpret_cast<int*>(ws + p.mq_ind), h_att, query, pos, neg, b0.hx,
                b0.xpg, (__half*)nullptr, sp, g.B, g.N, g.H, g.M, g.d);
  }
  // ---- decoder: the whole stack once per step (:184-192, :109-112) ----
  {
    CellW w = dec_w(g, p, ws);
    for (int t = 0; t < g.T_out; ++t) {
      CellBufs b = dec_bufs(g, p, ws, t);
      const float* go_src = nullptr;
      if (t > 0) go_src = (tf && tf[t - 1]) ? labels : output;
      MCRN_LAUNCH(k_stage_decoder_input, ew_grid((int64_t)g.R * g.Cdec), 256, 0, st, go_src, y_cov, const_cast<float*>(b.xpin), g.B,
                  g.T_out, g.N, g.Cout, g.Ycov, t, sp);
      MCRN_TRY(propagate_in(g, S, const_cast<float*>(b.xpin), b.xp_k, b.xp_n, g.B * g.Cdec, st));
      const bool last = (t + 1 == g.T_out);
      float* h_out = last ? ws + p.h_dec_last : dec_bufs(g, p, ws, t + 1).hx;
      float* h_mma = last ? nullptr : dec_bufs(g, p, ws, t + 1).xpg;
      MCRN_TRY(cell_forward(g, S, w, b, h_out, h_mma, st, nullptr));
      for (int l = 1; l < g.L; ++l) {
        const UpperW uw{ws + p.up[l - 1].d_wg, ws + p.up[l - 1].d_wu, g.D};
        UpperBufs ub = upper_dec_bufs(g, p, ws, l, t);
        MCRN_TRY(cell_forward_wide(g, S, uw, ub, dec_layer_out(g, p, ws, l - 1, t), ws + p.up_zh,
                                   ws + p.up[l - 1].d_hseq + p.dec_v_sz * (size_t)t, st));
      }
      MCRN_LAUNCH(k_proj_fwd, (int)ceil_div64(g.R, 8), 256, 0, st, dec_layer_out(g, p, ws, g.L - 1, t), prm->proj_w, prm->proj_b,
                  output, g.B, g.T_out, g.N, g.D, g.Cout, t);
    }
  }
  return MCRN_OK;
}

int backward_impl_layers(const Geo& g, const Plan& p, const mcrn_params* prm, const mcrn_layer_params* up, const uint8_t* tf,
                         const float* d_output, const float* d_hatt, const float* d_query, const float* d_pos,
                         const float* d_neg, const mcrn_params* grads, const mcrn_layer_params* ugrads, float* ws,
                         cudaStream_t st) {
  PerStageScope per_stage;
  const float* S = ws + p.Sr;
  MCRN_TRY(side_init());
  MCRN_TRY(fw_init());
  MCRN_CUDA_OK(cudaMemsetAsync(ws + p.acc_begin, 0, (p.acc_end - p.acc_begin) * sizeof(float), st));
  MCRN_CUDA_OK(cudaMemsetAsync(grads->memory, 0, (size_t)g.M * g.d * sizeof(float), st));
  MCRN_CUDA_OK(cudaMemsetAsync(grads->proj_w, 0, (size_t)g.Cout * g.D * sizeof(float), st));
  MCRN_CUDA_OK(cudaMemsetAsync(grads->proj_b, 0, (size_t)g.Cout * sizeof(float), st));
  const int64_t nD = g.R * g.D, nHh = g.R * g.H;
  float *dH0 = ws + p.dH, *dXin = ws + p.dXin, *dx = ws + p.up_dx;
  auto dec_dH = [&](int l) { return l == 0 ? dH0 : ws + p.up[l - 1].d_dH; };     // recurrent gradient of decoder layer l
  for (int l = 0; l < g.L; ++l) MCRN_CUDA_OK(cudaMemsetAsync(dec_dH(l), 0, (size_t)nD * sizeof(float), st));
  // ---- decoder, reverse time; within a step top layer first ----
  {
    CellW w = dec_w(g, p, ws);
    bool have_dgo = false;
    for (int t = g.T_out - 1; t >= 0; --t) {
      const bool use_dgo = have_dgo && !(tf && tf[t]);
      MCRN_LAUNCH(k_proj_bwd, (int)ceil_div64(g.R, 32), 256, (size_t)32 * g.Cout * sizeof(float), st, d_output,
                  use_dgo ? dXin : nullptr, g.Cdec, dec_layer_out(g, p, ws, g.L - 1, t), prm->proj_w, dec_dH(g.L - 1), 0,
                  grads->proj_w, grads->proj_b, g.B, g.T_out, g.N, g.D, g.Cout, t);
      for (int l = g.L - 1; l >= 1; --l) {
        const UpperPlan& u = p.up[l - 1];
        const UpperW uw{ws + u.d_wg, ws + u.d_wu, g.D};
        MCRN_TRY(cell_backward_wide(g, p, ws, S, uw, upper_dec_bufs(g, p, ws, l, t), ws + u.d_dU + (int64_t)t * nD,
                                    ws + u.d_dG + (int64_t)t * 2 * nD, dec_dH(l), dx, st));
        MCRN_LAUNCH(k_add_inplace, ew_grid(nD), 256, 0, st, dec_dH(l - 1), (const float*)dx, nD);
      }
      const bool need_dxin = (t > 0) && !(tf && tf[t - 1]);
      MCRN_TRY(cell_backward(g, p, ws, S, w, dec_bufs(g, p, ws, t), ws + p.d_dU + (int64_t)t * nD, ws + p.d_dG + (int64_t)t * 2 * nD,
                             dH0, dH0, need_dxin ? dXin : nullptr, st));
      have_dgo = need_dxin;
    }
    MCRN_TRY(acc_dw_all(g, ws + p.dec_xpu, (int64_t)p.dec_xp_sz, g.T_out, g.D, ws + p.d_dU, g.D, ws + p.a_d_wu, st));
    MCRN_TRY(acc_dw_all(g, ws + p.dec_xpg, (int64_t)p.dec_xp_sz, g.T_out, g.D, ws + p.d_dG, 2 * g.D, ws + p.a_d_wg, st));
    for (int l = 1; l < g.L; ++l) {
      const UpperPlan& u = p.up[l - 1];
      MCRN_TRY(acc_dw_all(g, ws + u.d_xpu, (int64_t)p.up_d_xp_sz, g.T_out, 2 * g.D, ws + u.d_dU, g.D, ws + u.a_d_wu, st));
      MCRN_TRY(acc_dw_all(g, ws + u.d_xpg, (int64_t)p.up_d_xp_sz, g.T_out, 2 * g.D, ws + u.d_dG, 2 * g.D, ws + u.a_d_wg, st));
      // every layer started from the same state (:181): the gradients of the initial states add up
      MCRN_LAUNCH(k_add_inplace, ew_grid(nD), 256, 0, st, dH0, (const float*)(ws + u.d_dH), nD);
    }
  }
  // ---- memory query (single stream) ----
  const float* h_top = enc_layer_out(g, p, ws, g.L - 1, g.T_in - 1);
  {
    size_t shm = (8 * (g.d + g.M) + (size_t)g.M * (g.d + 1)) * sizeof(float);
    float *dv = ws + p.mq_dv, *dsc = ws + p.mq_dsc, *dq = ws + p.mq_dq;
    MCRN_LAUNCH(k_memory_query_bwd_rows, (int)ceil_div64(g.R, 8), 256, shm, st, dH0, d_hatt, d_query, d_pos, d_neg,
                prm->memory, ws + p.mq_att, reinterpret_cast<const int*>(ws + p.mq_ind), dv, dsc, dq, grads->memory,
                g.B, g.N, g.H, g.M, g.d);
    const int sp = split_for(1, g.R / 16);
    {  // dMemory += att^T dv + dsc^T query
      GemmDesc q;
      q.A = ws + p.mq_att; q.a_row = 1; q.a_k = g.M; q.M = g.M; q.Kseg = (int)g.R;
      q.B = dv; q.b_k = g.d; q.b_n = 1; q.N = g.d; q.splits = sp; q.prec_exact = 1;
      EpiAtomicAdd e{grads->memory, g.d, 0};
      MCRN_TRY(gemm(q, e, st));
      q.A = dsc; q.B = ws + p.mq_q;
      MCRN_TRY(gemm(q, e, st));
    }
    {  // dWq = h_top^T dq
      MCRN_CUDA_OK(cudaMemsetAsync(grads->wq, 0, (size_t)g.H * g.d * sizeof(float), st));
      GemmDesc q;
      q.A = h_top; q.a_row = 1; q.a_k = g.H; q.M = g.H; q.Kseg = (int)g.R;
      q.B = dq; q.b_k = g.d; q.b_n = 1; q.N = g.d; q.splits = sp; q.prec_exact = 1;
      EpiAtomicAdd e{grads->wq, g.d, 0};
      MCRN_TRY(gemm(q, e, st));
    }
    {  // d(h_top) = dH0[:, :H] + dq Wq^T
      GemmDesc q;
      q.A = dq; q.a_row = g.d; q.a_k = 1; q.M = (int)g.R; q.Kseg = g.d;
      q.B = prm->wq; q.b_k = 1; q.b_n = g.d; q.N = g.H; q.prec_exact = 1;
      EpiStoreStrideAdd e{ws + p.dHenc, g.H, dH0, g.D};
      MCRN_TRY(gemm(q, e, st));
    }
  }
  // ---- encoder: top layer first, each layer in reverse time; dxseq[t] carries the gradient w.r.t. the outputs of the
  // layer below (written by layer l, consumed and overwritten in place by layer l-1) ----
  {
    float* dxseq = ws + p.up_dxseq;
    float* dHe = ws + p.dHenc;                       // top layer: gradient of its last state from the memory query
    for (int l = g.L - 1; l >= 1; --l) {
      const UpperPlan& u = p.up[l - 1];
      const UpperW uw{ws + u.e_wg, ws + u.e_wu, g.H};
      for (int t = g.T_in - 1; t >= 0; --t) {
        float* dxt = dxseq + p.enc_v_sz * (size_t)t;
        if (l < g.L - 1) MCRN_LAUNCH(k_add_inplace, ew_grid(nHh), 256, 0, st, dHe, (const float*)dxt, nHh);
        MCRN_TRY(cell_backward_wide(g, p, ws, S, uw, upper_enc_bufs(g, p, ws, l, t), ws + u.e_dU + (int64_t)t * nHh,
                                    ws + u.e_dG + (int64_t)t * 2 * nHh, dHe, dxt, st));
      }
      MCRN_TRY(acc_dw_all(g, ws + u.e_xpu, (int64_t)p.up_e_xp_sz, g.T_in, 2 * g.H, ws + u.e_dU, g.H, ws + u.a_e_wu, st));
      MCRN_TRY(acc_dw_all(g, ws + u.e_xpg, (int64_t)p.up_e_xp_sz, g.T_in, 2 * g.H, ws + u.e_dG, 2 * g.H, ws + u.a_e_wg, st));
      dHe = ws + p.up_dHe2;                          // the layers below receive gradient through dxseq only
      MCRN_CUDA_OK(cudaMemsetAsync(dHe, 0, (size_t)nHh * sizeof(float), st));
    }
    CellW w = enc_w(g, p, ws);
    for (int t = g.T_in - 1; t >= 0; --t) {
      if (g.L > 1) MCRN_LAUNCH(k_add_inplace, ew_grid(nHh), 256, 0, st, dHe, (const float*)(dxseq + p.enc_v_sz * (size_t)t), nHh);
      MCRN_TRY(cell_backward(g, p, ws, S, w, enc_bufs(g, p, ws, t), ws + p.e_dU + (int64_t)t * nHh, ws + p.e_dG + (int64_t)t * 2 * nHh,
                             dHe, dHe, nullptr, st));
    }
    MCRN_TRY(acc_dw_all(g, ws + p.enc_xpu, (int64_t)p.enc_xp_sz, g.T_in, g.H, ws + p.e_dU, g.H, ws + p.a_e_wu, st));
    MCRN_TRY(acc_dw_all(g, ws + p.enc_xpg, (int64_t)p.enc_xp_sz, g.T_in, g.H, ws + p.e_dG, 2 * g.H, ws + p.a_e_wg, st));
  }
  MCRN_TRY(side_join(st));        // the layer-0 dS accumulations (side stream of cell_backward) have landed
  MCRN_TRY(supports_backward(g, p, ws, prm, grads, st));
  // ---- un-fold the weight gradients into the reference layout ----
  MCRN_LAUNCH(k_unfold_grads, 128, 256, 0, st, ws + p.a_e_wg, grads->enc_gate_w, grads->enc_gate_b, g.Cin, g.H, 2 * g.H, g.cheb_k);
  MCRN_LAUNCH(k_unfold_grads, 128, 256, 0, st, ws + p.a_e_wu, grads->enc_update_w, grads->enc_update_b, g.Cin, g.H, g.H, g.cheb_k);
  MCRN_LAUNCH(k_unfold_grads, 128, 256, 0, st, ws + p.a_d_wg, grads->dec_gate_w, grads->dec_gate_b, g.Cdec, g.D, 2 * g.D, g.cheb_k);
  MCRN_LAUNCH(k_unfold_grads, 128, 256, 0, st, ws + p.a_d_wu, grads->dec_update_w, grads->dec_update_b, g.Cdec, g.D, g.D, g.cheb_k);
  for (int l = 1; l < g.L; ++l) {
    const UpperPlan& u = p.up[l - 1];
    const mcrn_layer_params& q = ugrads[l - 1];
    MCRN_LAUNCH(k_unfold_grads, 128, 256, 0, st, ws + u.a_e_wg, q.enc_gate_w, q.enc_gate_b, 0, 2 * g.H, 2 * g.H, g.cheb_k);
    MCRN_LAUNCH(k_unfold_grads, 128, 256, 0, st, ws + u.a_e_wu, q.enc_update_w, q.enc_update_b, 0, 2 * g.H, g.H, g.cheb_k);
    MCRN_LAUNCH(k_unfold_grads, 128, 256, 0, st, ws + u.a_d_wg, q.dec_gate_w, q.dec_gate_b, 0, 2 * g.D, 2 * g.D, g.cheb_k);
    MCRN_LAUNCH(k_unfold_grads, 128, 256, 0, st, ws + u.a_d_wu, q.dec_update_w, q.dec_update_b, 0, 2 * g.D, g.D, g.cheb_k);
  }
  return MCRN_OK;
}

