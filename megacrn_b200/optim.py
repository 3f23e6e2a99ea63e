"""The optimiser half of the reference training step, fused:

    torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)      model/traintest_MegaCRN.py:129
    torch.optim.Adam(model.parameters(), lr, eps=...).step()               :104, :130

as two kernels over the 14 parameter tensors (``mcrn_adam_step``; 14 + 8 per stacked layer: ``mcrn_adam_step_layers``):
a squared-norm reduction and the clipped Adam update.
Step count, learning rate, gradient norm and clip coefficient live in a 4-float device tensor, so the call needs no
host synchronisation and can be captured in the same CUDA graph as forward + loss + backward.
"""
from __future__ import annotations

import torch

from . import _abi


class FusedClipAdam:
    """Drop-in for ``clip_grad_norm_`` + ``torch.optim.Adam(...).step()`` on a ``megacrn_b200.MegaCRN``.

    ``lr`` can be changed between steps (``opt.lr = ...``, what MultiStepLR does at its milestones,
    model/traintest_MegaCRN.py:105, :132); ``last_grad_norm`` / ``last_clip_coef`` read the device scalars back."""

    def __init__(self, model, lr=0.01, betas=(0.9, 0.999), eps=1e-3, max_grad_norm=5.0):
        self.model = model
        sd = dict(model.named_parameters())
        self.params = [sd[k] for k in _abi.param_keys(getattr(model, "num_layers", 1))]      # 14, + 8 per stacked layer
        dev = self.params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedClipAdam needs the model on a CUDA device (megacrn_b200 has no CPU path)")
        self.betas, self.eps, self.max_grad_norm = betas, eps, max_grad_norm
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.state = torch.zeros(4, device=dev, dtype=torch.float32)      # step, lr, grad norm, clip coef
        self.state[1] = lr
        self._lr = lr

    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        self._lr = float(value)
        self.state[1:2].fill_(self._lr)

    @property
    def last_grad_norm(self):
        return float(self.state[2].sqrt())

    @property
    def last_clip_coef(self):
        return float(self.state[3])

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        lib = _abi.load()
        grads = []
        for p in self.params:
            if p.grad is None:
                raise RuntimeError("FusedClipAdam.step(): every parameter needs a gradient (run backward first)")
            grads.append(p.grad if p.grad.is_contiguous() else p.grad.contiguous())
        m = self.model
        layers = getattr(m, "num_layers", 1)
        dims = _abi.Dims(batch=1, num_nodes=m.num_nodes, seq_len=1, horizon=m.horizon, input_dim=m.input_dim,
                         output_dim=m.output_dim, ycov_dim=m.ycov_dim, rnn_units=m.rnn_units, num_layers=layers,
                         cheb_k=m.cheb_k, mem_num=m.mem_num, mem_dim=m.mem_dim)
        dev = self.params[0].device
        sets = (self.params, grads, self.exp_avg, self.exp_avg_sq)
        base = [_abi.make_params(t[:14]) for t in sets]
        with torch.cuda.device(dev):
            if layers == 1:
                st = lib.mcrn_adam_step(dims, base[0], base[1], base[2], base[3],
                                        self.state.data_ptr(), self.betas[0], self.betas[1], self.eps,
                                        self.max_grad_norm if self.max_grad_norm else 0.0,
                                        torch.cuda.current_stream(dev).cuda_stream)
            else:
                upper = [_abi.make_layer_params(t[14:]) for t in sets]
                st = lib.mcrn_adam_step_layers(dims, base[0], _abi.layer_ptr(upper[0]), base[1], _abi.layer_ptr(upper[1]),
                                               base[2], _abi.layer_ptr(upper[2]), base[3], _abi.layer_ptr(upper[3]),
                                               self.state.data_ptr(), self.betas[0], self.betas[1], self.eps,
                                               self.max_grad_norm if self.max_grad_norm else 0.0,
                                               torch.cuda.current_stream(dev).cuda_stream)
        _abi.check(st, "mcrn_adam_step")
        self.note_update()

    def note_update(self):
        """The update went through raw pointers: tell the module (eval fast-path cache) and torch's version counters.
        Also called by GraphedTrainStep after replaying a graph that contains this optimiser."""
        self.model.note_parameter_update()
        for p in self.params:
            torch.autograd.graph.increment_version(p)

    def state_dict(self):
        return {"exp_avg": [t.clone() for t in self.exp_avg], "exp_avg_sq": [t.clone() for t in self.exp_avg_sq],
                "state": self.state.clone(), "lr": self._lr}

    def load_state_dict(self, sd):
        for d, s in zip(self.exp_avg, sd["exp_avg"]):
            d.copy_(s)
        for d, s in zip(self.exp_avg_sq, sd["exp_avg_sq"]):
            d.copy_(s)
        self.state.copy_(sd["state"])
        self._lr = sd["lr"]
