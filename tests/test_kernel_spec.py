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


def test_fold_unfold_roundtrip():
    w = torch.arange(6 * 5 * 3, dtype=torch.float64).reshape(30, 3)
    st, win = K.fold_agcn_weights(w, 2, 3, 3)
    assert st.shape == (5, 3, 3) and win.shape == (5, 2, 3)
    g = K.unfold_agcn_grads(st, win, 2, 3, 3)
    wv, gv = w.reshape(6, 5, 3), g.reshape(6, 5, 3)
    assert torch.equal(gv[0], wv[0] + wv[3]) and torch.equal(gv[3], gv[0])
    assert torch.equal(gv[1], wv[1]) and torch.equal(gv[5], wv[5])
