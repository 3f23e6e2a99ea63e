"""CPU, float64: the restructured algebra + hand-derived backward == autograd of the
reference formulation (oracle).  This is the proof the CUDA decomposition rests on."""
import pytest
import torch

import kernel_spec as K
from oracle import megacrn_oracle as O


def _case(cheb_k, input_dim=1, seed=0):
    d = O.Dims(num_nodes=17, horizon=4, rnn_units=8, mem_num=5, mem_dim=12, cheb_k=cheb_k, input_dim=input_dim)
    p = {k: v.double() for k, v in O.init_params(d, seed=seed).items()}
    # non-zero biases so their gradients/paths are exercised
    g = torch.Generator().manual_seed(5)
    for k in p:
        if k.endswith("bias"):
            p[k] = torch.randn(p[k].shape, generator=g, dtype=torch.float64) * 0.1
    x, y_cov, labels = O.synthetic_batch(d, 3, 5, seed=99, dtype=torch.float64)
    return d, p, x, y_cov, labels


@pytest.mark.parametrize("cheb_k", [2, 3, 4])
@pytest.mark.parametrize("tf", [[False] * 4, [True, False, True, False], [True] * 4])
def test_forward_and_backward_match_autograd(cheb_k, tf):
    d, p, x, y_cov, labels = _case(cheb_k, input_dim=1 + (cheb_k == 4))
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ref = O.forward(d, q, x, y_cov, labels, tf)
    with torch.no_grad():
        got, saved = K.model_fwd(d, p, x, y_cov, labels, tf)
    for a, b in zip(got, ref):
        assert torch.allclose(a, b.detach(), rtol=1e-10, atol=1e-12)
    gen = torch.Generator().manual_seed(3)
    ups = [torch.randn(r.shape, generator=gen, dtype=torch.float64) for r in ref]
    loss = sum((r * u).sum() for r, u in zip(ref, ups))
    auto = torch.autograd.grad(loss, list(q.values()))
    with torch.no_grad():
        mine = K.model_bwd(d, p, saved, *ups)
    for (name, _), ga in zip(q.items(), auto):
        assert mine[name].shape == ga.shape, name
        err = (mine[name] - ga).abs().max().item()
        scale = ga.abs().max().item() + 1e-30
        assert err <= 1e-9 * scale + 1e-12, (name, err, scale)


@pytest.mark.parametrize("layers,cheb_k", [(2, 3), (3, 2)])
@pytest.mark.parametrize("tf", [[False] * 4, [True, False, True, False]])
def test_stacked_cells_match_autograd(layers, cheb_k, tf):
    """num_layers > 1 (model/MegaCRN.py:71-78, :109-112): the wide-operand cells of layers >= 1 ([x_in | h] as one AGCN operand,
    weights folded with cin = 0) and the layer-wise BPTT of csrc/model.cu against autograd of the reference formulation."""
    d = O.Dims(num_nodes=13, horizon=4, rnn_units=8, mem_num=5, mem_dim=12, cheb_k=cheb_k, num_layers=layers)
    p = {k: v.double() for k, v in O.init_params(d, seed=2).items()}
    g = torch.Generator().manual_seed(5)
    for k in p:
        if k.endswith("bias"):
            p[k] = torch.randn(p[k].shape, generator=g, dtype=torch.float64) * 0.1
    x, y_cov, labels = O.synthetic_batch(d, 3, 5, seed=98, dtype=torch.float64)
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ref = O.forward(d, q, x, y_cov, labels, tf)
    with torch.no_grad():
        got, saved = K.model_fwd_layers(d, p, x, y_cov, labels, tf)
    for a, b in zip(got, ref):
        assert torch.allclose(a, b.detach(), rtol=1e-10, atol=1e-12)
    gen = torch.Generator().manual_seed(3)
    ups = [torch.randn(r.shape, generator=gen, dtype=torch.float64) for r in ref]
    auto = torch.autograd.grad(sum((r * u).sum() for r, u in zip(ref, ups)), list(q.values()))
    with torch.no_grad():
        mine = K.model_bwd_layers(d, p, saved, *ups)
    assert set(mine) == set(q)
    for (name, _), ga in zip(q.items(), auto):
        assert mine[name].shape == ga.shape, name
        err = (mine[name] - ga).abs().max().item()
        scale = ga.abs().max().item() + 1e-30
        assert err <= 1e-9 * scale + 1e-12, (name, err, scale)


def test_fold_unfold_roundtrip():
    w = torch.arange(6 * 5 * 3, dtype=torch.float64).reshape(30, 3)
    st, win = K.fold_agcn_weights(w, 2, 3, 3)
    assert st.shape == (5, 3, 3) and win.shape == (5, 2, 3)
    g = K.unfold_agcn_grads(st, win, 2, 3, 3)
    wv, gv = w.reshape(6, 5, 3), g.reshape(6, 5, 3)
    assert torch.equal(gv[0], wv[0] + wv[3]) and torch.equal(gv[3], gv[0])
    assert torch.equal(gv[1], wv[1]) and torch.equal(gv[5], wv[5])


@pytest.mark.parametrize("nhalf", [1, 2])
def test_fused_backward_reassociation_matches_autograd(nhalf):
    """The algebra of csrc/agcn_bwd_fused_h.cuh, agcn_ds_fused_h.cuh and agcn_dw_fused_h.cuh (fp64, CPU): for one AGCN
    out = X W_0 + sum_k (S_k X) W_k  with upstream dV,
        dX   = dV W_0^T + sum_k (S_k^T dV) W_k^T            (Q_k = S_k^T dV, the block kept in tensor memory)
        dW_0 = X^T dV ,  dW_k = X^T Q_k                     (weight gradient from the stored Q blocks, not from P_k = S_k X)
        dS_k = sum_b (dV_b W_k^T) X_b^T                     (dXP = dV W_k^T only in tensor memory)
    and, for the 2H-wide gate AGCN, the column halves of dV can be propagated separately (Q_k = [Q_k,0 | Q_k,1])."""
    g = torch.Generator().manual_seed(7)
    KS, N, B, C = 4, 11, 3, 6
    O_ = nhalf * C
    S = torch.randn(KS, N, N, generator=g, dtype=torch.float64, requires_grad=True)
    X = torch.randn(N, B, C, generator=g, dtype=torch.float64, requires_grad=True)
    W = torch.randn(KS + 1, C, O_, generator=g, dtype=torch.float64, requires_grad=True)
    dV = torch.randn(N, B, O_, generator=g, dtype=torch.float64)
    out = X @ W[0] + sum(torch.einsum("nm,mbc->nbc", S[k], X) @ W[1 + k] for k in range(KS))
    gX, gW, gS = torch.autograd.grad((out * dV).sum(), [X, W, S])
    with torch.no_grad():
        # Q_k per column half, as the kernel propagates them
        Q = torch.stack([torch.cat([torch.einsum("nm,nbo->mbo", S[k], dV[..., h * C:(h + 1) * C]) for h in range(nhalf)], -1)
                         for k in range(KS)])
        dX = dV @ W[0].T + sum(Q[k] @ W[1 + k].T for k in range(KS))
        dW0 = torch.einsum("nbc,nbo->co", X, dV)
        dWk = torch.stack([torch.einsum("nbc,nbo->co", X, Q[k]) for k in range(KS)])
        dS = torch.stack([torch.einsum("nbc,mbc->nm", dV @ W[1 + k].T, X) for k in range(KS)])
    assert torch.allclose(dX, gX, rtol=1e-10, atol=1e-12)
    assert torch.allclose(dW0, gW[0], rtol=1e-10, atol=1e-12)
    assert torch.allclose(dWk, gW[1:], rtol=1e-10, atol=1e-12)
    assert torch.allclose(dS, gS, rtol=1e-10, atol=1e-12)


def test_loss_scale_is_exact_in_fp16():
    """A power-of-two loss scale commutes with fp16 rounding (no mantissa change) while the value stays in the normal range:
    fp16(x * s) / s == fp16-grid value of x -- the property the fp16 backward's fp32 outputs rely on."""
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4096, generator=g, dtype=torch.float32) * 3e-6            # gradient-sized values: subnormal in fp16 as they are
    s = 2.0 ** 24
    y = (x * s).to(torch.float16).to(torch.float32) / s
    rel = ((y - x).abs() / x.abs().clamp_min(1e-30)).max().item()
    assert rel <= 2.0 ** -11 * 1.0001                                        # 11-bit significand, round to nearest
    assert ((x.to(torch.float16).to(torch.float32) - x).abs() / x.abs().clamp_min(1e-30)).max().item() > 1e-2   # unscaled: lost
