"""GPU tests of the step that bench.py times and a trainer runs: ``GraphedTrainStep`` (CUDA-graph replay of forward +
fused trainer loss + backward [+ fused clip + Adam]) against the eager ``train_step`` and against the CPU oracle, the
eval fast path after raw-pointer weight updates, the host-buffer C entry, and the state rules of the
autograd function.  Reference: model/traintest_MegaCRN.py:114-130 (step), :50-99 (evaluate)."""
import ctypes as C

import numpy as np
import pytest
import torch

from golden_util import OUT_NAMES, rel_l2
from oracle import megacrn_oracle as O

pytestmark = pytest.mark.gpu

LOSS_KW = dict(scaler_mean=54.0, scaler_std=20.0)


def _dev():
    return torch.device("cuda:0")


def _model(d, p):
    from megacrn_b200 import MegaCRN
    m = MegaCRN(d.num_nodes, d.input_dim, d.output_dim, d.horizon, d.rnn_units, num_layers=d.num_layers,
                cheb_k=d.cheb_k, ycov_dim=d.ycov_dim, mem_num=d.mem_num, mem_dim=d.mem_dim,
                cl_decay_steps=d.cl_decay_steps, use_curriculum_learning=d.use_curriculum_learning).to(_dev())
    m.load_state_dict(p)
    return m


@pytest.fixture(autouse=True)
def _default_engine():
    from megacrn_b200 import _abi
    lib = _abi.load()
    prev = lib.mcrn_get_engine()
    lib.mcrn_set_engine(0)
    yield
    lib.mcrn_set_engine(prev)


@pytest.mark.parametrize("flags_kind", ["teacher_forced", "mixed", "free_running"])
def test_graphed_step_equals_eager_step_and_oracle_c2(flags_kind):
    """The thing bench.py times (graph replay, side streams, static gradient buffers, device-resident loss scale) gives
    the loss and the 14 gradients of the eager step, and both match the CPU oracle at the full C2 shape (B=64)."""
    from megacrn_b200.train_step import GraphedTrainStep, train_step
    d, B, T = O.Dims(num_nodes=207), 64, 12
    p = O.init_params(d, seed=0)
    x, y_cov, labels = O.synthetic_batch(d, B, T, seed=1234)
    flags = {"teacher_forced": [True] * 12, "mixed": [True, False] * 6, "free_running": [False] * 12}[flags_kind]
    dv = _dev()
    m = _model(d, p).train()
    # eager
    loss_e = train_step(m, x.to(dv), y_cov.to(dv), labels.to(dv), teacher_forcing=flags, **LOSS_KW)
    grads_e = {n: q.grad.detach().clone() for n, q in m.named_parameters()}
    loss_e = float(loss_e.item())
    # graph replay (twice: the second replay must not accumulate into the first one's gradients)
    g = GraphedTrainStep(m, B, T, **LOSS_KW)
    g.load(x, y_cov, labels)
    for _ in range(2):
        loss_g = float(g(teacher_forcing=flags).item())
    torch.cuda.synchronize()
    assert abs(loss_g - loss_e) <= 1e-5 * abs(loss_e), (loss_g, loss_e)
    for n, q in m.named_parameters():
        # same kernels, same operands; only the order of the split-K atomics differs
        assert rel_l2(q.grad.cpu(), grads_e[n].cpu()) < 2e-5, (n, rel_l2(q.grad.cpu(), grads_e[n].cpu()))
    assert g.flat_grad() is not None            # one flat buffer -> one all-reduce
    if flags_kind != "mixed":
        return                                   # one oracle run (minutes of CPU at B=64) is enough
    ref_loss, ref_outs, ref_grads = O.loss_and_grads(d, p, x, y_cov, labels, flags, **LOSS_KW)
    assert abs(loss_g - float(ref_loss)) <= 1e-3 * abs(float(ref_loss)), (loss_g, float(ref_loss))
    for n, q in m.named_parameters():
        # the step's OWN upstream gradient: sign(|.|) of residuals that straddle zero flips within rounding noise, which
        # moves the small supports-path gradients (We1 / We2) by ~1e-2; the tight bound under the reference's upstream
        # gradients is test_gpu_parity.py::test_full_size_vs_oracle
        assert rel_l2(q.grad.cpu(), ref_grads[n]) < 3e-2, (n, rel_l2(q.grad.cpu(), ref_grads[n]))


@pytest.mark.parametrize("layers", [1, 2])
def test_adam_trajectory_graphed_vs_torch_adam_on_oracle(layers):
    """5 steps of GraphedTrainStep(optimizer=FusedClipAdam) against clip_grad_norm_(5) + torch.optim.Adam(lr .01, eps 1e-3)
    driving the CPU oracle (model/traintest_MegaCRN.py:104, :128-130).  layers = 2: stacked cells -- 22 tensors in one
    gradient norm and one update (mcrn_adam_step_layers), the per-stage engine captured in the graph."""
    from megacrn_b200.optim import FusedClipAdam
    from megacrn_b200.train_step import GraphedTrainStep
    d, B, T = O.Dims(num_nodes=60, horizon=4, rnn_units=64, mem_num=8, mem_dim=64, num_layers=layers), 8, 4
    p = O.init_params(d, seed=4)
    flags = [True, False, True, True]
    batches = [O.synthetic_batch(d, B, T, seed=100 + i) for i in range(5)]
    # oracle trajectory
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    opt = torch.optim.Adam(list(q.values()), lr=0.01, eps=1e-3)
    ref_losses = []
    for x, y_cov, labels in batches:
        opt.zero_grad()
        loss = O.trainer_loss(O.forward(d, q, x, y_cov, labels, flags), labels, **LOSS_KW)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(q.values()), 5.0)
        opt.step()
        ref_losses.append(float(loss))
    m = _model(d, p).train()
    fopt = FusedClipAdam(m, lr=0.01, eps=1e-3, max_grad_norm=5.0)
    g = GraphedTrainStep(m, B, T, optimizer=fopt, **LOSS_KW)
    got = []
    for x, y_cov, labels in batches:
        g.load(x, y_cov, labels)
        got.append(float(g(teacher_forcing=flags).item()))
    for a, b in zip(got, ref_losses):
        assert abs(a - b) <= 2e-3 * abs(b), (got, ref_losses)
    sd = m.state_dict()
    for k, v in q.items():
        assert rel_l2(sd[k].cpu(), v.detach()) < 1e-2, (k, rel_l2(sd[k].cpu(), v.detach()))


@pytest.mark.parametrize("graphed", [False, True])
def test_eval_after_fused_adam_sees_the_new_weights(graphed):
    """FusedClipAdam updates the parameters through raw pointers (no tensor version bump, no Python at
    all under graph replay); the eval fast path must not reuse supports / folded weights cached before the update."""
    from megacrn_b200 import MegaCRN
    from megacrn_b200.optim import FusedClipAdam
    from megacrn_b200.train_step import GraphedTrainStep, train_step
    d, B, T = O.Dims(num_nodes=50, horizon=3, rnn_units=64, mem_num=8, mem_dim=64), 4, 3
    p = O.init_params(d, seed=6)
    x, y_cov, labels = O.synthetic_batch(d, B, T, seed=9)
    dv = _dev()
    xs, ys, ls = x.to(dv), y_cov.to(dv), labels.to(dv)
    m = _model(d, p)
    m.eval()
    with torch.no_grad():
        before = [o.clone() for o in m(xs, ys)]
        m(xs, ys)                                # second call: served from the cached prologue
    m.train()
    opt = FusedClipAdam(m, lr=0.01, eps=1e-3, max_grad_norm=5.0)
    flags = [True, False, True]
    if graphed:
        g = GraphedTrainStep(m, B, T, optimizer=opt, **LOSS_KW)
        g.load(x, y_cov, labels)
        for _ in range(3):
            g(teacher_forcing=flags)
    else:
        for _ in range(3):
            opt.zero_grad()
            train_step(m, xs, ys, ls, teacher_forcing=flags, **LOSS_KW)
            opt.step()
    m.eval()
    with torch.no_grad():
        after = m(xs, ys)
        fresh_model = _model(d, {k: v.detach().cpu() for k, v in m.state_dict().items()}).eval()
        fresh = fresh_model(xs, ys)
    assert rel_l2(after[0].cpu(), before[0].cpu()) > 1e-3          # the weights did move
    for k, a, b in zip(OUT_NAMES, after, fresh):
        assert torch.equal(a, b), k
    # an update made while the module STAYS in eval mode (no train()/eval() toggle in between)
    with torch.no_grad():
        m(xs, ys)
    for q in m.parameters():
        q.grad = torch.ones_like(q)
    opt.step()
    with torch.no_grad():
        after2 = m(xs, ys)
        fresh2 = _model(d, {k: v.detach().cpu() for k, v in m.state_dict().items()}).eval()(xs, ys)
    for k, a, b in zip(OUT_NAMES, after2, fresh2):
        assert torch.equal(a, b), k


def test_second_backward_on_the_same_forward_raises():
    """the first backward releases the workspace; a second one must fail loudly, not read overwritten
    activations."""
    d = O.Dims(num_nodes=40, horizon=2, rnn_units=64, mem_num=6, mem_dim=64)
    p = O.init_params(d, seed=2)
    x, y_cov, labels = O.synthetic_batch(d, 2, 2, seed=5)
    dv = _dev()
    m = _model(d, p).train()
    outs = m(x.to(dv), y_cov.to(dv), labels.to(dv), teacher_forcing=[True, False])
    loss = outs[0].sum() + outs[2].sum()
    loss.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="second time"):
        loss.backward()


def test_forward_host_entry_matches_module():
    """mcrn_forward_host (include/megacrn_b200.h: host buffers in, host buffers out -- the entry a non-PyTorch host binds)."""
    from megacrn_b200 import _abi
    lib = _abi.load()
    d, B, T = O.Dims(num_nodes=70, horizon=3, rnn_units=64, mem_num=10, mem_dim=64), 3, 4
    p = O.init_params(d, seed=8)
    x, y_cov, labels = O.synthetic_batch(d, B, T, seed=3)
    flags = [True, False, False]
    dv = _dev()
    m = _model(d, p).train()
    with torch.no_grad():
        want = [o.cpu() for o in m(x.to(dv), y_cov.to(dv), labels.to(dv), teacher_forcing=flags)]
    dims = _abi.Dims(batch=B, num_nodes=d.num_nodes, seq_len=T, horizon=d.horizon, input_dim=d.input_dim,
                     output_dim=d.output_dim, ycov_dim=d.ycov_dim, rnn_units=d.rnn_units, num_layers=1, cheb_k=d.cheb_k,
                     mem_num=d.mem_num, mem_dim=d.mem_dim)
    host = [p[k].contiguous().numpy() for k in _abi.STATE_DICT_KEYS]
    prm = _abi.Params()
    for name, a in zip(_abi.PARAM_FIELDS, host):
        setattr(prm, name, a.ctypes.data)
    nbytes = lib.mcrn_host_workspace_bytes(dims, 0)
    assert nbytes > 0
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dv)
    xn, yn, ln = x.numpy(), y_cov.numpy(), labels.numpy()
    out = np.full((B, d.horizon, d.num_nodes, d.output_dim), np.nan, np.float32)
    aux = [np.full((B, d.num_nodes, d.mem_dim), np.nan, np.float32) for _ in range(4)]
    st = lib.mcrn_forward_host(dims, prm, xn.ctypes.data, yn.ctypes.data, ln.ctypes.data, _abi.tf_bytes(flags, d.horizon),
                               out.ctypes.data, aux[0].ctypes.data, aux[1].ctypes.data, aux[2].ctypes.data,
                               aux[3].ctypes.data, ws.data_ptr(), nbytes, 0, torch.cuda.current_stream().cuda_stream)
    assert st == 0, lib.mcrn_last_error()
    for k, a, b in zip(OUT_NAMES, [out] + aux, want):
        assert np.array_equal(a, b.numpy()), k


def test_mask_count_and_dp_loss_normaliser():
    """mcrn_mask_count counts the labels whose inverse-scaled value is non-zero (model/utils.py:127); mcrn_trainer_loss_dp with
    that count reproduces mcrn_trainer_loss, and with another normaliser rescales exactly the masked-MAE term and d(output)."""
    from megacrn_b200 import _abi
    from megacrn_b200.train_step import fused_trainer_loss
    lib = _abi.load()
    d = O.Dims(num_nodes=30, horizon=4, rnn_units=64, mem_num=6, mem_dim=64)
    g = torch.Generator().manual_seed(2)
    B = 5
    out = torch.randn(B, d.horizon, d.num_nodes, 1, generator=g)
    lab = torch.randn(B, d.horizon, d.num_nodes, 1, generator=g)
    lab[1, :, :7] = -2.0                     # (-2) * 25 + 50 == 0 exactly -> masked
    qy, ps, ng = (torch.randn(B, d.num_nodes, d.mem_dim, generator=g) for _ in range(3))
    dv = _dev()
    kw = dict(scaler_mean=50.0, scaler_std=25.0)
    cnt = torch.zeros(1, device=dv)
    st = lib.mcrn_mask_count(lab.to(dv).data_ptr(), lab.numel(), 50.0, 25.0, cnt.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert st == 0, lib.mcrn_last_error()
    n_valid = int((lab * 25.0 + 50.0 != 0).sum())
    assert int(cnt.item()) == n_valid == lab.numel() - d.horizon * 7
    args = (d, out.to(dv), lab.to(dv), qy.to(dv), ps.to(dv), ng.to(dv))
    l0, do0, dq0 = fused_trainer_loss(*args, **kw)
    l1, do1, dq1 = fused_trainer_loss(*args, mask_count=cnt, **kw)
    assert torch.equal(l0, l1) and torch.equal(do0, do1) and torch.equal(dq0, dq1)
    ref = O.trainer_loss((out, None, qy, ps, ng), lab, **kw)
    assert abs(float(l0) - float(ref)) < 1e-5 * abs(float(ref))
    l2, do2, dq2 = fused_trainer_loss(*args, mask_count=cnt * 2, **kw)
    assert rel_l2(do2.cpu(), 0.5 * do0.cpu()) < 1e-6 and torch.equal(dq2, dq0)
    mae = float(O.masked_mae_loss(out * 25.0 + 50.0, lab * 25.0 + 50.0))
    assert abs((float(l0) - float(l2)) - 0.5 * mae) < 1e-4 * mae
