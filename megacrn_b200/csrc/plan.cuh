// Workspace layout ("plan") of one forward(+backward) of the MegaCRN hot path.
// All buffers live in ONE caller-provided device allocation; offsets are a pure function of
// the dims and the save-for-backward flag, so forward and backward agree without any state.
#pragma once

#include "common.cuh"

namespace mcrn {

// Buffers of one stacked cell pair (encoder + decoder cell of layer l >= 1, model/MegaCRN.py:62-63, :100-101).  Their AGCN
// operand is V = [x_in | h] of width 2*Hs (x_in = the state of the layer below), so every XP block is [R][2*Hs].
struct UpperPlan {
  size_t e_wg, e_wu, d_wg, d_wu;                 // folded weights [hi|lo][NB+1][2Hs][O] (cin = 0: block NB carries the bias row only)
  size_t e_xpg, e_xpu, e_z, e_r, e_hc, e_hseq;   // per slot; e_hseq: the layer's output of EVERY step [T_in][R][H] (input of the layer above)
  size_t d_xpg, d_xpu, d_z, d_r, d_hc, d_hseq;   // d_hseq [T_out][R][D]
  size_t e_dU, e_dG, d_dU, d_dG, d_dH;           // backward: dU / dG of every step, the layer's recurrent decoder gradient [R][D]
  size_t a_e_wg, a_e_wu, a_d_wg, a_d_wu;         // weight-gradient accumulators [NB+1][2Hs][O] (inside the zeroed region)
};

struct Plan {
  Geo g;
  bool save;
  int enc_slots, dec_slots;
  size_t bytes;
  // ---- offsets in floats --------------------------------------------------
  size_t S, Sr, E1, E2, L1, L2;        // Sr: TF32-rounded copy of the supports (tensor-core operand)
  size_t e_wg, e_wu, d_wg, d_wu;         // folded weights [hi|lo][NB+1][Hs][O] (block NB = input channels + bias)
  size_t enc_xpin;                       // [NB][N][T_in][B][Cin]
  size_t enc_xpg, enc_xpu, enc_z, enc_r, enc_hc, enc_hx;   // per slot; hx = exact fp32 input state of the step
  size_t h_enc;                          // [R][H]
  size_t mq_q, mq_att, mq_ind;           // [R][d], [R][M], int[R][2]
  size_t dec_xpin, dec_xpg, dec_xpu, dec_z, dec_r, dec_hc, dec_hx;
  size_t h_dec_last;                     // [R][D]
  // fp16 operand copies of the fused forward (agcn_fused_h.cuh); offsets in floats, contents are halves
  size_t s16, e_wg16, e_wu16, d_wg16, d_wu16;
  size_t enc_x16, enc_zh16, enc_ib16;     // row-major [R][Hs]; read as K-major A and as MN-major B operands
  size_t dec_x16, dec_zh16, dec_ib16;
  size_t enc_ib16c, dec_ib16c;           // compact input blocks: [T_in][R][64] / [T_out][R][64] halves
  size_t enc_ib32c, dec_ib32c;           // [T][R][16] floats (training: operand of the input-block weight gradient)
  // per-slot sizes (floats)
  size_t enc_xp_sz, enc_v_sz, dec_xpin_sz, dec_xp_sz, dec_v_sz;
  // ---- backward temporaries -------------------------------------------------
  size_t dH, dXP, dXP2, dIBu, dHp, dXPin, dXin, mq_dv, mq_dsc, mq_dq, dHenc;   // dXP: update-AGCN blocks, dXP2: gate-AGCN blocks
  size_t e_dU, e_dG, d_dU, d_dG;         // dU_t / dG_t of every step ([T][R][Hs] / [T][R][2Hs]): weight gradients are one GEMM per AGCN over all steps
  size_t acc_begin, acc_end;             // zeroed at the start of backward
  size_t dS, a_e_wg, a_e_wu, a_d_wg, a_d_wu;
  size_t dg1, dg2;
  size_t dLa, dLb, dL1, dE1, dE2;
  // fused backward (agcn_bwd_fused.cuh): transposed supports, Q blocks of every step, input-block gradients,
  // per-step dXPin
  size_t St, e_Qu, e_Qg, d_Qu, d_Qg, dIBu16, dIBg16, dXPin_all, dHr;
  // fp16 fused backward (agcn_bwd_fused_h.cuh): loss scale {s, 1/s, amax bits}, transposed fp16 supports, fp16 weights,
  // scaled fp16 operand copies of dU / dG (row-major and node-transposed)
  size_t gs, s16T, e_wg16n, e_wu16n, d_wg16n, d_wu16n, dU16, dG16, e_dU16, e_dG16;
  // fp16 weight-gradient kernel (agcn_dw_fused_h.cuh): row-major fp16 Q blocks of every step
  size_t e_Qu16T, e_Qg16T, d_Qu16T, d_Qg16T;
  size_t dXPin_sz;
  // loss scratch
  size_t loss_scratch;                   // 8 floats
  // ---- stacked layers (num_layers > 1): all zero-sized when num_layers == 1 ----
  UpperPlan up[MCRN_MAX_LAYERS - 1];
  int up_slots_e, up_slots_d;            // slots of the per-step buffers (all steps when saving, else 1)
  size_t up_e_xp_sz, up_d_xp_sz;         // floats per XP slot: (NB+1)*R*2H / (NB+1)*R*2D
  size_t up_zero, up_zh;                 // [R][D] zeros (initial encoder state of a stacked layer); z*h scratch [R][D]
  size_t up_dXP, up_dXP2, up_dV0, up_dhp, up_dxa, up_dx;   // backward temporaries of a stacked cell
  size_t up_dxseq, up_dHe2;              // [T_in][R][H] gradient w.r.t. the outputs of the layer below; [R][H] recurrent encoder gradient
};

static inline int make_plan(const Geo& g, bool save, Plan* p) {
  memset(p, 0, sizeof(*p));
  p->g = g;
  p->save = save;
  size_t off = 0;
  auto take = [&](size_t nfloats) {
    size_t o = off;
    off += (nfloats + 63) / 64 * 64;      // 256-byte granularity keeps every buffer TMA/vector aligned
    return o;
  };
  const size_t R = (size_t)g.R, N = g.N, ldS = g.ldS, NB = g.NB, KS = g.KS;
  p->S = take(KS * N * ldS);
  p->Sr = take(KS * N * ldS);
  p->E1 = take(N * g.d);
  p->E2 = take(N * g.d);
  p->L1 = take(N * ldS);
  p->L2 = take(N * ldS);
  p->e_wg = take(2 * (NB + 1) * g.H * 2 * g.H);
  p->e_wu = take(2 * (NB + 1) * g.H * g.H);
  p->d_wg = take(2 * (NB + 1) * g.D * 2 * g.D);
  p->d_wu = take(2 * (NB + 1) * g.D * g.D);
  p->enc_slots = (save || g.L > 1) ? g.T_in : 2;      // a stacked layer reads the outputs of every step of the layer below
  p->dec_slots = save ? g.T_out : 2;
  p->enc_xpin = take(NB * N * g.T_in * g.B * g.Cin);
  p->enc_xp_sz = ((NB + 1) * R * g.H + 63) / 64 * 64;      // NB state blocks + the input block
  p->enc_v_sz = (R * g.H + 63) / 64 * 64;
  p->enc_xpg = take(p->enc_xp_sz * p->enc_slots);
  p->enc_xpu = take(p->enc_xp_sz * p->enc_slots);
  p->enc_z = take(p->enc_v_sz * p->enc_slots);
  p->enc_r = take(p->enc_v_sz * p->enc_slots);
  p->enc_hc = take(p->enc_v_sz * p->enc_slots);
  p->enc_hx = take(p->enc_v_sz * p->enc_slots);
  p->h_enc = take(R * g.H);
  p->mq_q = take(R * g.d);
  p->mq_att = take(R * g.M);
  p->mq_ind = take(R * 2);
  p->dec_xpin_sz = (NB * R * g.Cdec + 63) / 64 * 64;
  p->dec_xp_sz = ((NB + 1) * R * g.D + 63) / 64 * 64;
  p->dec_v_sz = (R * g.D + 63) / 64 * 64;
  p->dec_xpin = take(p->dec_xpin_sz * p->dec_slots);
  p->dec_xpg = take(p->dec_xp_sz * p->dec_slots);
  p->dec_xpu = take(p->dec_xp_sz * p->dec_slots);
  p->dec_z = take(p->dec_v_sz * p->dec_slots);
  p->dec_r = take(p->dec_v_sz * p->dec_slots);
  p->dec_hc = take(p->dec_v_sz * p->dec_slots);
  p->dec_hx = take(p->dec_v_sz * p->dec_slots);
  p->h_dec_last = take(R * g.D);
  {
    const size_t ld16 = (N + 7) / 8 * 8;
    auto halves = [&](size_t n) { return take((n + 1) / 2); };
    p->s16 = halves(KS * N * ld16);
    p->e_wg16 = halves(2 * (NB + 1) * g.H * 2 * g.H);
    p->e_wu16 = halves(2 * (NB + 1) * g.H * g.H);
    p->d_wg16 = halves(2 * (NB + 1) * g.D * 2 * g.D);
    p->d_wu16 = halves(2 * (NB + 1) * g.D * g.D);
    // training: the row-major fp16 state copies of EVERY step are kept (operands of the fp16 support-gradient kernel)
    const size_t es = save ? (size_t)g.T_in : 1, ds = save ? (size_t)g.T_out : 1;
    p->enc_x16 = halves(es * R * g.H);
    p->enc_zh16 = halves(es * R * g.H);
    p->enc_ib16 = halves(R * g.H);
    p->dec_x16 = halves(ds * R * g.D);
    p->dec_zh16 = halves(ds * R * g.D);
    p->dec_ib16 = halves(R * g.D);
    p->enc_ib16c = halves((size_t)g.T_in * R * 64);
    p->dec_ib16c = halves((size_t)g.T_out * R * 64);
    p->enc_ib32c = take(save ? (size_t)g.T_in * R * 16 : 0);
    p->dec_ib32c = take(save ? (size_t)g.T_out * R * 16 : 0);
  }
  p->loss_scratch = take(64);
  if (g.L > 1) {
    const size_t kwe = 2 * (size_t)g.H, kwd = 2 * (size_t)g.D;
    p->up_slots_e = save ? g.T_in : 1;
    p->up_slots_d = save ? g.T_out : 1;
    p->up_e_xp_sz = ((NB + 1) * R * kwe + 63) / 64 * 64;
    p->up_d_xp_sz = ((NB + 1) * R * kwd + 63) / 64 * 64;
    p->up_zero = take(R * g.D);
    p->up_zh = take(R * g.D);
    for (int l = 0; l + 1 < g.L; ++l) {
      UpperPlan& u = p->up[l];
      u.e_wg = take(2 * (NB + 1) * kwe * 2 * g.H);
      u.e_wu = take(2 * (NB + 1) * kwe * g.H);
      u.d_wg = take(2 * (NB + 1) * kwd * 2 * g.D);
      u.d_wu = take(2 * (NB + 1) * kwd * g.D);
      u.e_xpg = take(p->up_e_xp_sz * p->up_slots_e);
      u.e_xpu = take(p->up_e_xp_sz * p->up_slots_e);
      u.e_z = take(p->enc_v_sz * p->up_slots_e);
      u.e_r = take(p->enc_v_sz * p->up_slots_e);
      u.e_hc = take(p->enc_v_sz * p->up_slots_e);
      u.e_hseq = take(p->enc_v_sz * (size_t)g.T_in);
      u.d_xpg = take(p->up_d_xp_sz * p->up_slots_d);
      u.d_xpu = take(p->up_d_xp_sz * p->up_slots_d);
      u.d_z = take(p->dec_v_sz * p->up_slots_d);
      u.d_r = take(p->dec_v_sz * p->up_slots_d);
      u.d_hc = take(p->dec_v_sz * p->up_slots_d);
      u.d_hseq = take(p->dec_v_sz * (size_t)g.T_out);
    }
  }
  if (save) {
    const size_t Cm = (size_t)(g.Cin > g.Cdec ? g.Cin : g.Cdec);
    p->dH = take(R * g.D);
    p->e_dU = take((size_t)g.T_in * R * g.H);
    p->e_dG = take((size_t)g.T_in * R * 2 * g.H);
    p->d_dU = take((size_t)g.T_out * R * g.D);
    p->d_dG = take((size_t)g.T_out * R * 2 * g.D);
    p->dXP = take((NB + 1) * R * g.D);
    p->dXP2 = take((NB + 1) * R * g.D);
    p->dIBu = take(R * g.D);
    p->dHp = take(R * g.D);
    p->dXPin = take(NB * R * Cm);
    p->dXin = take(R * Cm);
    p->mq_dv = take(R * g.d);
    p->mq_dsc = take(R * g.M);
    p->mq_dq = take(R * g.d);
    p->dHenc = take(R * g.H);
    p->St = take(KS * N * ldS);
    // Q blocks of every step: only where the fused backward is instantiated (hidden width 64 / 128)
    const bool eq = (g.H == 64 || g.H == 128), dq = (g.D == 64 || g.D == 128);
    p->e_Qu = take(eq ? (size_t)g.T_in * KS * R * g.H : 0);
    p->e_Qg = take(eq ? (size_t)g.T_in * 2 * KS * R * g.H : 0);
    p->d_Qu = take(dq ? (size_t)g.T_out * KS * R * g.D : 0);
    p->d_Qg = take(dq ? (size_t)g.T_out * 2 * KS * R * g.D : 0);
    p->dHr = take(R * g.D);
    {
      const size_t ld16 = (N + 7) / 8 * 8;
      auto halves = [&](size_t n) { return take((n + 1) / 2); };
      p->gs = take(64);
      p->s16T = halves(KS * N * ld16);
      p->e_wg16n = halves((NB + 1) * g.H * 2 * g.H);
      p->e_wu16n = halves((NB + 1) * g.H * g.H);
      p->d_wg16n = halves((NB + 1) * g.D * 2 * g.D);
      p->d_wu16n = halves((NB + 1) * g.D * g.D);
      // row-major copies of every step (fp16 support-gradient kernel); decoder and encoder separately: the decoder's
      // support-gradient kernels run on the side stream while the encoder backward proceeds
      p->dU16 = halves((size_t)g.T_out * R * g.D);
      p->dG16 = halves((size_t)g.T_out * R * 2 * g.D);
      p->e_dU16 = halves((size_t)g.T_in * R * g.H);
      p->e_dG16 = halves((size_t)g.T_in * R * 2 * g.H);
      // row-major fp16 Q blocks of every step (operand of the fp16 weight-gradient kernel)
      p->e_Qu16T = halves(eq ? (size_t)g.T_in * KS * R * g.H : 0);
      p->e_Qg16T = halves(eq ? (size_t)g.T_in * 2 * KS * R * g.H : 0);
      p->d_Qu16T = halves(dq ? (size_t)g.T_out * KS * R * g.D : 0);
      p->d_Qg16T = halves(dq ? (size_t)g.T_out * 2 * KS * R * g.D : 0);
    }
    p->dIBu16 = take(R * 16);
    p->dIBg16 = take(R * 16);
    p->dXPin_sz = (NB * R * Cm + 63) / 64 * 64;
    p->dXPin_all = take(p->dXPin_sz * (size_t)(g.T_in > g.T_out ? g.T_in : g.T_out));
    if (g.L > 1) {
      const size_t kwd = 2 * (size_t)g.D;
      p->up_dXP = take((NB + 1) * R * kwd);
      p->up_dXP2 = take((NB + 1) * R * kwd);
      p->up_dV0 = take(R * kwd);
      p->up_dhp = take(R * g.D);
      p->up_dxa = take(R * g.D);
      p->up_dx = take(R * g.D);
      p->up_dxseq = take(p->enc_v_sz * (size_t)g.T_in);
      p->up_dHe2 = take(R * g.H);
      for (int l = 0; l + 1 < g.L; ++l) {
        UpperPlan& u = p->up[l];
        u.e_dU = take((size_t)g.T_in * R * g.H);
        u.e_dG = take((size_t)g.T_in * R * 2 * g.H);
        u.d_dU = take((size_t)g.T_out * R * g.D);
        u.d_dG = take((size_t)g.T_out * R * 2 * g.D);
        u.d_dH = take(R * g.D);
      }
    }
    p->acc_begin = off;
    if (g.L > 1) {
      const size_t kwe = 2 * (size_t)g.H, kwd = 2 * (size_t)g.D;
      for (int l = 0; l + 1 < g.L; ++l) {
        UpperPlan& u = p->up[l];
        u.a_e_wg = take((NB + 1) * kwe * 2 * g.H);
        u.a_e_wu = take((NB + 1) * kwe * g.H);
        u.a_d_wg = take((NB + 1) * kwd * 2 * g.D);
        u.a_d_wu = take((NB + 1) * kwd * g.D);
      }
    }
    p->dS = take(KS * N * ldS);
    p->a_e_wg = take((NB + 1) * g.H * 2 * g.H);
    p->a_e_wu = take((NB + 1) * g.H * g.H);
    p->a_d_wg = take((NB + 1) * g.D * 2 * g.D);
    p->a_d_wu = take((NB + 1) * g.D * g.D);
    p->dg1 = take(N * ldS);
    p->dg2 = take(N * ldS);
    p->acc_end = off;
    p->dLa = take(N * ldS);
    p->dLb = take(N * ldS);
    p->dL1 = take(N * ldS);
    p->dE1 = take(N * g.d);
    p->dE2 = take(N * g.d);
  }
  p->bytes = off * sizeof(float);
  return MCRN_OK;
}

}  // namespace mcrn
