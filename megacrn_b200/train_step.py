"""The caller's side of one training step, fused: the trainer's loss
(model/traintest_MegaCRN.py:118-125) evaluated by mcrn_trainer_loss, with the analytic
d(loss)/d(output) and d(loss)/d(query) fed straight into the model's backward."""
from __future__ import annotations

import torch

from . import _abi


def _dims_from(d_like, batch, t_in=1):
    g = lambda *names: next(getattr(d_like, n) for n in names if hasattr(d_like, n))
    return _abi.Dims(batch=batch, num_nodes=g("num_nodes"), seq_len=t_in, horizon=g("horizon"),
                     input_dim=g("input_dim"), output_dim=g("output_dim"), ycov_dim=g("ycov_dim"),
                     rnn_units=g("rnn_units"), num_layers=1, cheb_k=g("cheb_k"), mem_num=g("mem_num"),
                     mem_dim=g("mem_dim"))


def fused_trainer_loss(d_like, output, labels, query, pos, neg, scaler_mean=54.0, scaler_std=20.0,
                       lamb=0.01, lamb1=0.01, want_grads=True):
    """Returns (loss[1], d_output, d_query); pos/neg are constants (the trainer detaches them)."""
    lib = _abi.load()
    dev = output.device
    dims = _dims_from(d_like, output.shape[0])
    c = lambda t: t.detach().to(torch.float32).contiguous()
    output, labels, query, pos, neg = map(c, (output, labels, query, pos, neg))
    loss = torch.empty(1, device=dev, dtype=torch.float32)
    d_out = torch.empty_like(output) if want_grads else None
    d_q = torch.empty_like(query) if want_grads else None
    scratch = torch.empty(64, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        st = lib.mcrn_trainer_loss(dims, output.data_ptr(), labels.data_ptr(), query.data_ptr(), pos.data_ptr(),
                                   neg.data_ptr(), scaler_mean, scaler_std, lamb, lamb1, loss.data_ptr(),
                                   _abi.ptr(d_out), _abi.ptr(d_q), scratch.data_ptr(), 256,
                                   torch.cuda.current_stream(dev).cuda_stream)
    _abi.check(st, "mcrn_trainer_loss")
    return loss, d_out, d_q


def train_step(model, x, y_cov, labels, batches_seen=0, teacher_forcing=None, **loss_kw):
    """forward + trainer loss + backward; leaves gradients in ``p.grad``; returns loss[1] (device)."""
    outs = model(x, y_cov, labels, batches_seen, teacher_forcing=teacher_forcing)
    output, _h_att, query, pos, neg = outs
    loss, d_out, d_q = fused_trainer_loss(model, output, labels, query, pos, neg, **loss_kw)
    torch.autograd.backward([output, query], [d_out, d_q])
    return loss
