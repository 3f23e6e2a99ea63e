"""Kernel-level specification of the B200 path (test infrastructure).

The same decomposition the CUDA library uses -- node-major state ``[N, B, H]``,
support set hoisted out of the time loop, the two identity blocks of every AGCN
weight folded into one, input channels separated from state channels, and a
hand-derived BPTT backward -- written with plain torch ops so that

  * the algebra is proven against the oracle (autograd of the reference
    formulation) on CPU, in float64, before any kernel is trusted;
  * each CUDA stage has a stage-level twin for the GPU unit tests.

Naming follows megacrn_b200/csrc: XP = "[X | P_1..P_KS]" block buffer, S = the KS
real supports [g1, T2(g1), .., g2, T2(g2), ..], KS = 2*(cheb_k-1).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch

Tensor = torch.Tensor


# ----------------------------------------------------------------------------
# parameter re-packing (done once per forward by mcrn "prep" kernels)
# ----------------------------------------------------------------------------
def fold_agcn_weights(w: Tensor, cin: int, hid: int, cheb_k: int):
    """w [2*cheb_k*(cin+hid), O] -> (Wst [1+KS, hid, O], Win [1+KS, cin, O]).

    Reference row index = k*(cin+hid) + c (model/MegaCRN.py:24-27); blocks k=0 and
    k=cheb_k multiply the identity support (:20) and are summed into block 0."""
    c = cin + hid
    wv = w.reshape(2 * cheb_k, c, -1)
    blocks = [wv[0] + wv[cheb_k]]
    for g in range(2):
        for k in range(1, cheb_k):
            blocks.append(wv[g * cheb_k + k])
    f = torch.stack(blocks, 0)                      # [1+KS, c, O]
    return f[:, cin:, :].contiguous(), f[:, :cin, :].contiguous()


def unfold_agcn_grads(d_wst: Tensor, d_win: Tensor, cin: int, hid: int, cheb_k: int) -> Tensor:
    """Inverse of fold_agcn_weights for gradients: both identity blocks receive block 0."""
    f = torch.cat([d_win, d_wst], dim=1)            # [1+KS, c, O]
    out = []
    idx = 1
    for g in range(2):
        out.append(f[0])
        for k in range(1, cheb_k):
            out.append(f[idx])
            idx += 1
    return torch.cat(out, dim=0)


def supports_fwd(p: Dict[str, Tensor], cheb_k: int):
    """model/MegaCRN.py:169-173 + the Chebyshev recursion of :19-23, hoisted."""
    mem = p["memory.Memory"]
    e1 = p["memory.We1"] @ mem
    e2 = p["memory.We2"] @ mem
    l1 = e1 @ e2.T
    g1 = torch.softmax(torch.relu(l1), -1)
    g2 = torch.softmax(torch.relu(l1.T), -1)
    n = g1.shape[0]
    eye = torch.eye(n, dtype=g1.dtype, device=g1.device)
    s_list = []
    for g in (g1, g2):
        ks = [eye, g]
        for _ in range(2, cheb_k):
            ks.append(2 * g @ ks[-1] - ks[-2])
        s_list.extend(ks[1:])
    saved = dict(e1=e1, e2=e2, l1=l1, g1=g1, g2=g2)
    return torch.stack(s_list, 0), saved           # [KS, N, N]


def supports_bwd(p, cheb_k, saved, d_s: Tensor):
    """d_s [KS,N,N] -> grads of We1, We2 and the Memory contribution."""
    mem = p["memory.Memory"]
    e1, e2, l1, g1, g2 = (saved[k] for k in ("e1", "e2", "l1", "g1", "g2"))
    n = g1.shape[0]
    eye = torch.eye(n, dtype=g1.dtype, device=g1.device)
    per = cheb_k - 1
    d_g = []
    for gi, g in enumerate((g1, g2)):
        ks = [eye, g]
        for _ in range(2, cheb_k):
            ks.append(2 * g @ ks[-1] - ks[-2])
        d_t = [torch.zeros_like(g) for _ in range(cheb_k)]
        for k in range(1, cheb_k):
            d_t[k] = d_s[gi * per + (k - 1)].clone()
        dg = torch.zeros_like(g)
        for k in range(cheb_k - 1, 1, -1):          # T_k = 2 g T_{k-1} - T_{k-2}
            dg = dg + 2 * d_t[k] @ ks[k - 1].T
            d_t[k - 1] = d_t[k - 1] + 2 * g.T @ d_t[k]
            d_t[k - 2] = d_t[k - 2] - d_t[k]
        dg = dg + d_t[1]
        d_g.append(dg)
    # softmax + relu backward; logits of g2 are the transpose of those of g1
    def sm_bwd(g, dg):
        return g * (dg - (g * dg).sum(-1, keepdim=True))
    dl_a = sm_bwd(g1, d_g[0]) * (l1 > 0)
    dl_b = sm_bwd(g2, d_g[1]) * (l1.T > 0)
    dl1 = dl_a + dl_b.T
    de1 = dl1 @ e2
    de2 = dl1.T @ e1
    return dict(We1=de1 @ mem.T, We2=de2 @ mem.T,
                Memory=p["memory.We1"].T @ de1 + p["memory.We2"].T @ de2)


def propagate(s: Tensor, x_nm: Tensor) -> Tensor:
    """[KS,N,N] x [N,B,C] -> [KS,N,B,C]  (model/MegaCRN.py:24-25, one GEMM of M=KS*N)."""
    ks, n, _ = s.shape
    _, b, c = x_nm.shape
    return (s.reshape(ks * n, n) @ x_nm.reshape(n, b * c)).reshape(ks, n, b, c)


def propagate_t(s: Tensor, d_p: Tensor) -> Tensor:
    """Adjoint of propagate w.r.t. x: sum_k S_k^T dP_k -> [N,B,C]."""
    ks, n, b, c = d_p.shape
    return (s.reshape(ks * n, n).T @ d_p.reshape(ks * n, b * c)).reshape(n, b, c)


def d_supports(d_p: Tensor, x_nm: Tensor) -> Tensor:
    """Adjoint of propagate w.r.t. S: dS_k = dP_k x^T -> [KS,N,N]."""
    ks, n, b, c = d_p.shape
    return (d_p.reshape(ks * n, b * c) @ x_nm.reshape(n, b * c).T).reshape(ks, n, n)


# ----------------------------------------------------------------------------
# one cell, forward and backward
# ----------------------------------------------------------------------------
def cell_fwd(s, h, xin, wg_st, wg_in, bg, wu_st, wu_in, bu):
    """h [N,B,H], xin [N,B,Cin]; returns h' and everything the backward reads."""
    hid = h.shape[-1]
    xp_in = torch.cat([xin[None], propagate(s, xin)], 0)                 # [1+KS,N,B,Cin]
    xp_g = torch.cat([h[None], propagate(s, h)], 0)                      # [1+KS,N,B,H]
    g = torch.einsum("knbc,kco->nbo", xp_g, wg_st) + torch.einsum("knbc,kco->nbo", xp_in, wg_in) + bg
    zr = torch.sigmoid(g)
    z, r = zr[..., :hid], zr[..., hid:]
    zh = z * h
    xp_u = torch.cat([zh[None], propagate(s, zh)], 0)
    u = torch.einsum("knbc,kco->nbo", xp_u, wu_st) + torch.einsum("knbc,kco->nbo", xp_in, wu_in) + bu
    hc = torch.tanh(u)
    h_new = r * h + (1 - r) * hc
    return h_new, dict(xp_in=xp_in, xp_g=xp_g, xp_u=xp_u, z=z, r=r, hc=hc)


def cell_bwd(s, sv, d_hnew, wg_st, wg_in, wu_st, wu_in, acc):
    """Returns (d_h, d_xin) and accumulates parameter/support grads into ``acc``."""
    h = sv["xp_g"][0]
    z, r, hc = sv["z"], sv["r"], sv["hc"]
    d_u = d_hnew * (1 - r) * (1 - hc * hc)
    d_r = d_hnew * (h - hc)
    d_h = d_hnew * r
    # update AGCN
    acc["wu_st"] += torch.einsum("knbc,nbo->kco", sv["xp_u"], d_u)
    acc["wu_in"] += torch.einsum("knbc,nbo->kco", sv["xp_in"], d_u)
    acc["bu"] += d_u.sum((0, 1))
    d_xp_u = torch.einsum("nbo,kco->knbc", d_u, wu_st)
    d_xp_in = torch.einsum("nbo,kco->knbc", d_u, wu_in)
    d_zh = d_xp_u[0] + propagate_t(s, d_xp_u[1:])
    acc["s"] += d_supports(d_xp_u[1:], sv["xp_u"][0])
    d_z = d_zh * h
    d_h = d_h + d_zh * z
    d_g = torch.cat([d_z * z * (1 - z), d_r * r * (1 - r)], -1)
    # gate AGCN
    acc["wg_st"] += torch.einsum("knbc,nbo->kco", sv["xp_g"], d_g)
    acc["wg_in"] += torch.einsum("knbc,nbo->kco", sv["xp_in"], d_g)
    acc["bg"] += d_g.sum((0, 1))
    d_xp_g = torch.einsum("nbo,kco->knbc", d_g, wg_st)
    d_xp_in = d_xp_in + torch.einsum("nbo,kco->knbc", d_g, wg_in)
    d_h = d_h + d_xp_g[0] + propagate_t(s, d_xp_g[1:])
    acc["s"] += d_supports(d_xp_g[1:], h)
    # input channels
    acc["s"] += d_supports(d_xp_in[1:], sv["xp_in"][0])
    d_xin = d_xp_in[0] + propagate_t(s, d_xp_in[1:])
    return d_h, d_xin


# ----------------------------------------------------------------------------
# memory query                                         model/MegaCRN.py:159-166
# ----------------------------------------------------------------------------
def memory_query_fwd(h_t: Tensor, mem: Tensor, wq: Tensor):
    """h_t [N,B,H] (node-major) -> value, query [N,B,d], att [N,B,M], ind [N,B,2]."""
    query = h_t @ wq
    att = torch.softmax(query @ mem.T, -1)
    value = att @ mem
    ind = torch.topk(att, 2, dim=-1).indices
    return value, query, att, ind


def memory_query_bwd(h_t, mem, wq, query, att, ind, d_value, d_query, d_pos, d_neg):
    d_att = d_value @ mem.T
    d_mem = torch.einsum("nbm,nbd->md", att, d_value)
    d_sc = att * (d_att - (att * d_att).sum(-1, keepdim=True))
    d_q = d_query + d_sc @ mem
    d_mem = d_mem + torch.einsum("nbm,nbd->md", d_sc, query)
    m, dd = mem.shape
    d_mem = d_mem.index_add(0, ind[..., 0].reshape(-1), d_pos.reshape(-1, dd))
    d_mem = d_mem.index_add(0, ind[..., 1].reshape(-1), d_neg.reshape(-1, dd))
    d_wq = torch.einsum("nbh,nbd->hd", h_t, d_q)
    d_h = d_q @ wq.T
    return d_h, d_mem, d_wq


# ----------------------------------------------------------------------------
# whole model (num_layers == 1)
# ----------------------------------------------------------------------------
def _nm(t: Tensor) -> Tensor:            # [B,N,C] -> node-major [N,B,C]
    return t.permute(1, 0, 2).contiguous()


def model_fwd(d, p: Dict[str, Tensor], x, y_cov, labels, tf: Sequence[bool]):
    assert d.num_layers == 1
    hid, dd, ck = d.rnn_units, d.decoder_dim, d.cheb_k
    s, s_saved = supports_fwd(p, ck)
    e = "encoder.dcrnn_cells.0."
    q = "decoder.dcrnn_cells.0."
    ew = fold_agcn_weights(p[e + "gate.weights"], d.input_dim, hid, ck) + \
        fold_agcn_weights(p[e + "update.weights"], d.input_dim, hid, ck)
    dw = fold_agcn_weights(p[q + "gate.weights"], d.output_dim + d.ycov_dim, dd, ck) + \
        fold_agcn_weights(p[q + "update.weights"], d.output_dim + d.ycov_dim, dd, ck)
    bsz, t_in = x.shape[0], x.shape[1]
    h = torch.zeros(d.num_nodes, bsz, hid, dtype=x.dtype, device=x.device)
    enc_saved = []
    for t in range(t_in):
        h, sv = cell_fwd(s, h, _nm(x[:, t]), ew[0], ew[1], p[e + "gate.bias"], ew[2], ew[3], p[e + "update.bias"])
        enc_saved.append(sv)
    h_enc = h
    value, query, att, ind = memory_query_fwd(h_enc, p["memory.Memory"], p["memory.Wq"])
    mem = p["memory.Memory"]
    pos, neg = mem[ind[..., 0]], mem[ind[..., 1]]
    h = torch.cat([h_enc, value], -1)
    go = torch.zeros(d.num_nodes, bsz, d.output_dim, dtype=x.dtype, device=x.device)
    wp, bp = p["proj.0.weight"], p["proj.0.bias"]
    dec_saved, outs, h_dec = [], [], []
    for t in range(d.horizon):
        xin = torch.cat([go, _nm(y_cov[:, t])], -1)
        h, sv = cell_fwd(s, h, xin, dw[0], dw[1], p[q + "gate.bias"], dw[2], dw[3], p[q + "update.bias"])
        dec_saved.append(sv)
        h_dec.append(h)
        go = h @ wp.T + bp
        outs.append(go)
        if tf[t]:
            go = _nm(labels[:, t])
    output = torch.stack([o.permute(1, 0, 2) for o in outs], 1)                 # [B,T,N,Cout]
    res = (output, value.permute(1, 0, 2), query.permute(1, 0, 2), pos.permute(1, 0, 2), neg.permute(1, 0, 2))
    saved = dict(s=s, s_saved=s_saved, ew=ew, dw=dw, enc=enc_saved, dec=dec_saved, h_dec=h_dec, h_enc=h_enc,
                 query=query, att=att, ind=ind, tf=list(tf))
    return res, saved


def model_bwd(d, p, saved, d_output, d_hatt, d_query, d_pos, d_neg) -> Dict[str, Tensor]:
    hid, dd, ck = d.rnn_units, d.decoder_dim, d.cheb_k
    cin_d = d.output_dim + d.ycov_dim
    s, ew, dw = saved["s"], saved["ew"], saved["dw"]
    wp = p["proj.0.weight"]
    z = torch.zeros_like

    def new_acc(w4):
        return dict(wg_st=z(w4[0]), wg_in=z(w4[1]), wu_st=z(w4[2]), wu_in=z(w4[3]),
                    bg=torch.zeros(w4[0].shape[-1], dtype=s.dtype, device=s.device),
                    bu=torch.zeros(w4[2].shape[-1], dtype=s.dtype, device=s.device), s=z(s))
    acc_d, acc_e = new_acc(dw), new_acc(ew)
    d_wp, d_bp = z(wp), z(p["proj.0.bias"])
    n, bsz = saved["h_enc"].shape[0], saved["h_enc"].shape[1]
    d_h = torch.zeros(n, bsz, dd, dtype=s.dtype, device=s.device)
    d_go_next = None                                   # grad w.r.t. the decoder input of step t+1
    for t in range(d.horizon - 1, -1, -1):
        d_out_t = _nm(d_output[:, t]).clone()
        if d_go_next is not None and not saved["tf"][t]:
            d_out_t = d_out_t + d_go_next               # go_{t+1} = out_t unless teacher-forced
        h_t = saved["h_dec"][t]
        d_wp += torch.einsum("nbo,nbd->od", d_out_t, h_t)
        d_bp += d_out_t.sum((0, 1))
        d_h = d_h + d_out_t @ wp
        d_h, d_xin = cell_bwd(s, saved["dec"][t], d_h, dw[0], dw[1], dw[2], dw[3], acc_d)
        d_go_next = d_xin[..., :d.output_dim]
    d_value = d_h[..., hid:] + _nm(d_hatt)
    d_henc, d_mem, d_wq = memory_query_bwd(saved["h_enc"], p["memory.Memory"], p["memory.Wq"], saved["query"],
                                           saved["att"], saved["ind"], d_value, _nm(d_query), _nm(d_pos), _nm(d_neg))
    d_h = d_h[..., :hid] + d_henc
    for t in range(len(saved["enc"]) - 1, -1, -1):
        d_h, _ = cell_bwd(s, saved["enc"][t], d_h, ew[0], ew[1], ew[2], ew[3], acc_e)
    sg = supports_bwd(p, ck, saved["s_saved"], acc_d["s"] + acc_e["s"])
    e, q = "encoder.dcrnn_cells.0.", "decoder.dcrnn_cells.0."
    return {
        "memory.Memory": d_mem + sg["Memory"], "memory.Wq": d_wq, "memory.We1": sg["We1"], "memory.We2": sg["We2"],
        e + "gate.weights": unfold_agcn_grads(acc_e["wg_st"], acc_e["wg_in"], d.input_dim, hid, ck),
        e + "gate.bias": acc_e["bg"],
        e + "update.weights": unfold_agcn_grads(acc_e["wu_st"], acc_e["wu_in"], d.input_dim, hid, ck),
        e + "update.bias": acc_e["bu"],
        q + "gate.weights": unfold_agcn_grads(acc_d["wg_st"], acc_d["wg_in"], cin_d, dd, ck),
        q + "gate.bias": acc_d["bg"],
        q + "update.weights": unfold_agcn_grads(acc_d["wu_st"], acc_d["wu_in"], cin_d, dd, ck),
        q + "update.bias": acc_d["bu"],
        "proj.0.weight": d_wp, "proj.0.bias": d_bp,
    }


# ----------------------------------------------------------------------------
# stacked cells (num_layers > 1)                 model/MegaCRN.py:62-63, :71-78, :100-101, :109-112
# ----------------------------------------------------------------------------
# The cells of layers >= 1 take the H- (D-) wide state of the layer below as their input.  The CUDA path runs them as an
# AGCN over the concatenated operand V = [x_in | h] of width 2*Hs with NO separate input channels: the reference's weight
# rows are already ordered (support block, [input channels | state channels]) (model/MegaCRN.py:42, :24-27), so
# fold_agcn_weights(w, cin=0, hid=2*Hs) is the folded weight of that operand and the bias rides in the input block.
def cell_fwd_wide(s, h, x_in, wg, bg, wu, bu):
    """h, x_in [N,B,Hs]; wg [1+KS, 2Hs, 2Hs], wu [1+KS, 2Hs, Hs] (folded with cin = 0)."""
    hid = h.shape[-1]
    v0g = torch.cat([x_in, h], -1)
    xp_g = torch.cat([v0g[None], propagate(s, v0g)], 0)                  # [1+KS,N,B,2Hs]
    zr = torch.sigmoid(torch.einsum("knbc,kco->nbo", xp_g, wg) + bg)
    z, r = zr[..., :hid], zr[..., hid:]
    v0u = torch.cat([x_in, z * h], -1)
    xp_u = torch.cat([v0u[None], propagate(s, v0u)], 0)
    hc = torch.tanh(torch.einsum("knbc,kco->nbo", xp_u, wu) + bu)
    h_new = r * h + (1 - r) * hc
    return h_new, dict(xp_g=xp_g, xp_u=xp_u, z=z, r=r, hc=hc, h=h)


def cell_bwd_wide(s, sv, d_hnew, wg, wu, acc):
    """Returns (d_h, d_x_in); accumulates wg / wu / bg / bu / s into ``acc``.  Step order = csrc/model.cu:cell_backward_wide."""
    h, z, r, hc = sv["h"], sv["z"], sv["r"], sv["hc"]
    hid = h.shape[-1]
    d_u = d_hnew * (1 - r) * (1 - hc * hc)
    acc["wu"] += torch.einsum("knbc,nbo->kco", sv["xp_u"], d_u)
    acc["bu"] += d_u.sum((0, 1))
    d_xp = torch.einsum("nbo,kco->knbc", d_u, wu)
    acc["s"] += d_supports(d_xp[1:], sv["xp_u"][0])
    d_v0 = d_xp[0] + propagate_t(s, d_xp[1:])                            # [N,B,2Hs]
    d_zh, d_xa = d_v0[..., hid:], d_v0[..., :hid]
    d_g = torch.cat([d_zh * h * z * (1 - z), d_hnew * (h - hc) * r * (1 - r)], -1)
    d_hp = d_hnew * r + d_zh * z
    acc["wg"] += torch.einsum("knbc,nbo->kco", sv["xp_g"], d_g)
    acc["bg"] += d_g.sum((0, 1))
    d_xp2 = torch.einsum("nbo,kco->knbc", d_g, wg)
    acc["s"] += d_supports(d_xp2[1:], sv["xp_g"][0])
    d_v0 = d_xp2[0] + propagate_t(s, d_xp2[1:])
    return d_hp + d_v0[..., hid:], d_xa + d_v0[..., :hid]


def model_fwd_layers(d, p: Dict[str, Tensor], x, y_cov, labels, tf: Sequence[bool]):
    """Layer-major encoder, step-major decoder: the data flow of csrc/model.cu:forward_impl_layers."""
    hid, dd, ck, L = d.rnn_units, d.decoder_dim, d.cheb_k, d.num_layers
    s, s_saved = supports_fwd(p, ck)
    cin_d = d.output_dim + d.ycov_dim
    e0, q0 = "encoder.dcrnn_cells.0.", "decoder.dcrnn_cells.0."
    ew = fold_agcn_weights(p[e0 + "gate.weights"], d.input_dim, hid, ck) + fold_agcn_weights(p[e0 + "update.weights"], d.input_dim, hid, ck)
    dw = fold_agcn_weights(p[q0 + "gate.weights"], cin_d, dd, ck) + fold_agcn_weights(p[q0 + "update.weights"], cin_d, dd, ck)
    ewl = [(fold_agcn_weights(p[f"encoder.dcrnn_cells.{i}.gate.weights"], 0, 2 * hid, ck)[0],
            fold_agcn_weights(p[f"encoder.dcrnn_cells.{i}.update.weights"], 0, 2 * hid, ck)[0]) for i in range(1, L)]
    dwl = [(fold_agcn_weights(p[f"decoder.dcrnn_cells.{i}.gate.weights"], 0, 2 * dd, ck)[0],
            fold_agcn_weights(p[f"decoder.dcrnn_cells.{i}.update.weights"], 0, 2 * dd, ck)[0]) for i in range(1, L)]
    bsz, t_in = x.shape[0], x.shape[1]
    kw = dict(dtype=x.dtype, device=x.device)
    # encoder, layer by layer (:71-78)
    seq = []
    h = torch.zeros(d.num_nodes, bsz, hid, **kw)
    enc_saved = [[]]
    for t in range(t_in):
        h, sv = cell_fwd(s, h, _nm(x[:, t]), ew[0], ew[1], p[e0 + "gate.bias"], ew[2], ew[3], p[e0 + "update.bias"])
        enc_saved[0].append(sv)
        seq.append(h)
    for i in range(1, L):
        pre = f"encoder.dcrnn_cells.{i}."
        h = torch.zeros(d.num_nodes, bsz, hid, **kw)
        nxt, svs = [], []
        for t in range(t_in):
            h, sv = cell_fwd_wide(s, h, seq[t], ewl[i - 1][0], p[pre + "gate.bias"], ewl[i - 1][1], p[pre + "update.bias"])
            svs.append(sv)
            nxt.append(h)
        enc_saved.append(svs)
        seq = nxt
    h_enc = seq[-1]
    value, query, att, ind = memory_query_fwd(h_enc, p["memory.Memory"], p["memory.Wq"])
    mem = p["memory.Memory"]
    pos, neg = mem[ind[..., 0]], mem[ind[..., 1]]
    states = [torch.cat([h_enc, value], -1)] * L                       # :179-181
    go = torch.zeros(d.num_nodes, bsz, d.output_dim, **kw)
    wp, bp = p["proj.0.weight"], p["proj.0.bias"]
    dec_saved, outs, h_top = [], [], []
    for t in range(d.horizon):
        xin = torch.cat([go, _nm(y_cov[:, t])], -1)
        new_states, svs = [], []
        h, sv = cell_fwd(s, states[0], xin, dw[0], dw[1], p[q0 + "gate.bias"], dw[2], dw[3], p[q0 + "update.bias"])
        new_states.append(h)
        svs.append(sv)
        for i in range(1, L):
            pre = f"decoder.dcrnn_cells.{i}."
            h, sv = cell_fwd_wide(s, states[i], h, dwl[i - 1][0], p[pre + "gate.bias"], dwl[i - 1][1], p[pre + "update.bias"])
            new_states.append(h)
            svs.append(sv)
        states = new_states
        dec_saved.append(svs)
        h_top.append(h)
        go = h @ wp.T + bp
        outs.append(go)
        if tf[t]:
            go = _nm(labels[:, t])
    output = torch.stack([o.permute(1, 0, 2) for o in outs], 1)
    res = (output, value.permute(1, 0, 2), query.permute(1, 0, 2), pos.permute(1, 0, 2), neg.permute(1, 0, 2))
    saved = dict(s=s, s_saved=s_saved, ew=ew, dw=dw, ewl=ewl, dwl=dwl, enc=enc_saved, dec=dec_saved, h_top=h_top, h_enc=h_enc,
                 query=query, att=att, ind=ind, tf=list(tf))
    return res, saved


def model_bwd_layers(d, p, saved, d_output, d_hatt, d_query, d_pos, d_neg) -> Dict[str, Tensor]:
    """BPTT over the stack: the data flow of csrc/model.cu:backward_impl_layers."""
    hid, dd, ck, L = d.rnn_units, d.decoder_dim, d.cheb_k, d.num_layers
    cin_d = d.output_dim + d.ycov_dim
    s, ew, dw, ewl, dwl = saved["s"], saved["ew"], saved["dw"], saved["ewl"], saved["dwl"]
    wp = p["proj.0.weight"]
    z = torch.zeros_like
    kw = dict(dtype=s.dtype, device=s.device)

    def acc0(w4):
        return dict(wg_st=z(w4[0]), wg_in=z(w4[1]), wu_st=z(w4[2]), wu_in=z(w4[3]), bg=torch.zeros(w4[0].shape[-1], **kw),
                    bu=torch.zeros(w4[2].shape[-1], **kw), s=z(s))

    def accw(w2):
        return dict(wg=z(w2[0]), wu=z(w2[1]), bg=torch.zeros(w2[0].shape[-1], **kw), bu=torch.zeros(w2[1].shape[-1], **kw), s=z(s))
    acc_d, acc_e = acc0(dw), acc0(ew)
    acc_dl, acc_el = [accw(w) for w in dwl], [accw(w) for w in ewl]
    d_wp, d_bp = z(wp), z(p["proj.0.bias"])
    n, bsz = saved["h_enc"].shape[0], saved["h_enc"].shape[1]
    d_hl = [torch.zeros(n, bsz, dd, **kw) for _ in range(L)]            # recurrent gradient of every decoder layer
    d_go_next = None
    for t in range(d.horizon - 1, -1, -1):
        d_out_t = _nm(d_output[:, t]).clone()
        if d_go_next is not None and not saved["tf"][t]:
            d_out_t = d_out_t + d_go_next
        d_wp += torch.einsum("nbo,nbd->od", d_out_t, saved["h_top"][t])
        d_bp += d_out_t.sum((0, 1))
        d_hl[L - 1] = d_hl[L - 1] + d_out_t @ wp
        for i in range(L - 1, 0, -1):
            d_hl[i], d_x = cell_bwd_wide(s, saved["dec"][t][i], d_hl[i], dwl[i - 1][0], dwl[i - 1][1], acc_dl[i - 1])
            d_hl[i - 1] = d_hl[i - 1] + d_x
        d_hl[0], d_xin = cell_bwd(s, saved["dec"][t][0], d_hl[0], dw[0], dw[1], dw[2], dw[3], acc_d)
        d_go_next = d_xin[..., :d.output_dim]
    d_h = sum(d_hl)                                                      # every layer started from the same state (:181)
    d_value = d_h[..., hid:] + _nm(d_hatt)
    d_henc, d_mem, d_wq = memory_query_bwd(saved["h_enc"], p["memory.Memory"], p["memory.Wq"], saved["query"],
                                           saved["att"], saved["ind"], d_value, _nm(d_query), _nm(d_pos), _nm(d_neg))
    t_in = len(saved["enc"][0])
    d_h = d_h[..., :hid] + d_henc                                        # gradient of the top layer's last state
    d_seq = [None] * t_in                                                # gradient w.r.t. the outputs of the layer below
    for i in range(L - 1, 0, -1):
        for t in range(t_in - 1, -1, -1):
            if d_seq[t] is not None:
                d_h = d_h + d_seq[t]
            d_h, d_seq[t] = cell_bwd_wide(s, saved["enc"][i][t], d_h, ewl[i - 1][0], ewl[i - 1][1], acc_el[i - 1])
        d_h = torch.zeros(n, bsz, hid, **kw)
    for t in range(t_in - 1, -1, -1):
        if d_seq[t] is not None:
            d_h = d_h + d_seq[t]
        d_h, _ = cell_bwd(s, saved["enc"][0][t], d_h, ew[0], ew[1], ew[2], ew[3], acc_e)
    d_s = acc_d["s"] + acc_e["s"] + sum(a["s"] for a in acc_dl) + sum(a["s"] for a in acc_el)
    sg = supports_bwd(p, ck, saved["s_saved"], d_s)
    e, q = "encoder.dcrnn_cells.0.", "decoder.dcrnn_cells.0."
    out = {
        "memory.Memory": d_mem + sg["Memory"], "memory.Wq": d_wq, "memory.We1": sg["We1"], "memory.We2": sg["We2"],
        e + "gate.weights": unfold_agcn_grads(acc_e["wg_st"], acc_e["wg_in"], d.input_dim, hid, ck), e + "gate.bias": acc_e["bg"],
        e + "update.weights": unfold_agcn_grads(acc_e["wu_st"], acc_e["wu_in"], d.input_dim, hid, ck), e + "update.bias": acc_e["bu"],
        q + "gate.weights": unfold_agcn_grads(acc_d["wg_st"], acc_d["wg_in"], cin_d, dd, ck), q + "gate.bias": acc_d["bg"],
        q + "update.weights": unfold_agcn_grads(acc_d["wu_st"], acc_d["wu_in"], cin_d, dd, ck), q + "update.bias": acc_d["bu"],
        "proj.0.weight": d_wp, "proj.0.bias": d_bp,
    }
    for i in range(1, L):
        for pre, a, w in ((f"encoder.dcrnn_cells.{i}.", acc_el[i - 1], 2 * hid), (f"decoder.dcrnn_cells.{i}.", acc_dl[i - 1], 2 * dd)):
            empty = a["wg"][:, :0, :]
            out[pre + "gate.weights"] = unfold_agcn_grads(a["wg"], empty, 0, w, ck)
            out[pre + "gate.bias"] = a["bg"]
            out[pre + "update.weights"] = unfold_agcn_grads(a["wu"], a["wu"][:, :0, :], 0, w, ck)
            out[pre + "update.bias"] = a["bu"]
    return out
