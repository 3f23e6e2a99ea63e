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


def fused_trainer_loss(d_like, output, labels, query, pos, neg, *, scaler_mean, scaler_std,
                       lamb=0.01, lamb1=0.01, want_grads=True, mask_count=None):
    """Returns (loss[1], d_output, d_query); pos/neg are constants (the trainer detaches them).

    ``scaler_mean`` / ``scaler_std``: the trainer's StandardScaler (model/traintest_MegaCRN.py:274-277, :118-119).
    ``mask_count``: 1-element device tensor holding the masked-MAE normaliser to use instead of this batch's own count
    of non-zero labels (data parallel: global count / world size, ``ddp.global_mask_count``)."""
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
        st = lib.mcrn_trainer_loss_dp(dims, output.data_ptr(), labels.data_ptr(), query.data_ptr(), pos.data_ptr(),
                                      neg.data_ptr(), scaler_mean, scaler_std, lamb, lamb1, _abi.ptr(mask_count),
                                      loss.data_ptr(), _abi.ptr(d_out), _abi.ptr(d_q), scratch.data_ptr(), 256,
                                      torch.cuda.current_stream(dev).cuda_stream)
    _abi.check(st, "mcrn_trainer_loss")
    return loss, d_out, d_q


DEFAULT_SCALER = dict(scaler_mean=54.0, scaler_std=20.0)     # the synthetic scaler of SURVEY.md 8(d); a trainer passes its own


def train_step(model, x, y_cov, labels, batches_seen=0, teacher_forcing=None, group=None, mask_count=None, **loss_kw):
    """forward + trainer loss + backward; leaves gradients in ``p.grad``; returns loss[1] (device).

    Inside an initialised process group with more than one rank the masked-MAE normaliser is the global one
    (``ddp.global_mask_count``: a 4-byte all-reduce issued before the forward and waited for only at the loss), so that
    the rank-average of loss and gradients equals the single-process step on the concatenated batch.  ``mask_count``: a
    1-element device tensor that already holds that normaliser (``GraphedTrainStep`` computes it before replaying its
    graph); no collective is issued here then."""
    from . import ddp
    for k, v in DEFAULT_SCALER.items():
        loss_kw.setdefault(k, v)
    pending = None
    if mask_count is None:
        pending = ddp.global_mask_count_begin(labels, loss_kw["scaler_mean"], loss_kw["scaler_std"], group)
    outs = model(x, y_cov, labels, batches_seen, teacher_forcing=teacher_forcing)
    output, _h_att, query, pos, neg = outs
    if pending is not None:
        mask_count = ddp.global_mask_count_end(pending)
    loss, d_out, d_q = fused_trainer_loss(model, output, labels, query, pos, neg, mask_count=mask_count, **loss_kw)
    torch.autograd.backward([output, query], [d_out, d_q])
    return loss


class GraphedTrainStep:
    """The training step (forward + fused trainer loss + backward) replayed from a CUDA graph.

    The step is ~500 small kernel launches; enqueueing them from the host costs several ms, more than the
    kernels need.  The whole step is captured once per teacher-forcing pattern (the coin flips of
    model/MegaCRN.py:188-191 are host control flow: they are still drawn from ``np.random`` on every call, in the
    reference's order, and select which captured graph is replayed) and replayed with one launch.

    Inputs are copied into static device buffers (``load``), gradients land in static ``p.grad`` tensors that
    alias one flat buffer (``flat_grad``), the loss in ``loss`` (a 1-element device tensor).
    """

    def __init__(self, model, batch, seq_len, max_graphs=16, optimizer=None, group=None, **loss_kw):
        """optimizer: a ``megacrn_b200.optim.FusedClipAdam``.  Single process: clip + Adam are part of the captured step.
        Inside a process group with more than one rank the graph holds forward + loss + backward only; every call then
        runs, around the replay: the 4-byte all-reduce of the masked-MAE normaliser (before), ONE gradient all-reduce
        (NCCL averages inside the collective) and the optimiser (after).  Collectives are not captured: graph capture of
        the NCCL calls hung on the 2-GPU box of this round (profiles/r2_summary.md)."""
        import torch.distributed as dist
        self.group = group
        self.dp = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.allreduce = self.dp            # the gradient all-reduce is issued by __call__
        self.model, self.loss_kw, self.max_graphs, self.optimizer = model, loss_kw, max_graphs, optimizer
        dev = next(model.parameters()).device
        self.x = torch.zeros(batch, seq_len, model.num_nodes, model.input_dim, device=dev)
        self.y_cov = torch.zeros(batch, model.horizon, model.num_nodes, model.ycov_dim, device=dev)
        self.labels = torch.zeros(batch, model.horizon, model.num_nodes, model.output_dim, device=dev)
        self.graphs = {}
        self.kernels_replayed = 0          # library kernels executed through graph replays so far
        self.loss = None
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.pool = None
        self.mask_count = torch.zeros(1, device=dev) if self.dp else None      # static: read by the captured loss kernel
        self._count_valid = False          # the normaliser depends on the labels only: recomputed after every load()

    def load(self, x, y_cov, labels, non_blocking=True):
        """Copy a batch into the static input buffers (the only supported way to change them: the data-parallel
        normaliser is recomputed when, and only when, new labels have been loaded)."""
        self.x.copy_(x, non_blocking=non_blocking)
        self.y_cov.copy_(y_cov, non_blocking=non_blocking)
        self.labels.copy_(labels, non_blocking=non_blocking)
        self._count_valid = False

    def _eager(self, flags, complete=True):
        for p in self.params:
            p.grad = None
        if complete:
            self._before()
        loss = self._step_body(flags)
        if complete:
            self._after()
        return loss

    def _step_body(self, flags):
        """What is captured: forward + loss + backward, and (single process only) the optimiser."""
        loss = train_step(self.model, self.x, self.y_cov, self.labels, teacher_forcing=flags, group=self.group,
                          mask_count=self.mask_count, **self.loss_kw)
        if self.optimizer is not None and not self.dp:
            self.optimizer.step()
        return loss

    def _before(self):
        if self.dp and not self._count_valid:
            self._count_valid = True
            from .ddp import global_mask_count_into
            kw = {**DEFAULT_SCALER, **self.loss_kw}
            global_mask_count_into(self.labels, kw["scaler_mean"], kw["scaler_std"], self.mask_count, self.group)

    def _after(self):
        if self.dp:
            from .ddp import allreduce_gradients
            allreduce_gradients(self.params, self.group)
            if self.optimizer is not None:
                self.optimizer.step()

    def _capture(self, flags):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        saved = None
        if self.optimizer is not None:                 # the warm-up steps must not advance the optimiser / the weights
            saved = (self.optimizer.state_dict(), [p.detach().clone() for p in self.params])
        self._before()                                 # (data parallel: a valid normaliser for the warm-up steps)
        with torch.cuda.stream(side):                  # warm-up outside capture (lazy inits, allocator); no collectives
            for _ in range(2):
                self._eager(flags, complete=False)
        torch.cuda.current_stream().wait_stream(side)      # the restore below must not overtake the warm-up updates
        if saved is not None:
            with torch.no_grad():
                self.optimizer.load_state_dict(saved[0])
                for p, q in zip(self.params, saved[1]):
                    p.copy_(q)
            self.optimizer.note_update()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        for p in self.params:
            p.grad = None
        lib = _abi.load()
        n0 = lib.mcrn_launch_count()
        with torch.cuda.graph(g, pool=self.pool):
            loss = self._step_body(flags)
        kernels = int(lib.mcrn_launch_count() - n0)       # library kernels recorded in this graph
        if self.pool is None:
            self.pool = g.pool()
        grads = [p.grad for p in self.params]
        return g, loss, grads, kernels

    def __call__(self, batches_seen=0, teacher_forcing=None):
        m = self.model
        flags = teacher_forcing if teacher_forcing is not None else m.draw_teacher_forcing(batches_seen)
        m.last_teacher_forcing = flags
        key = None if flags is None else tuple(bool(f) for f in flags)
        entry = self.graphs.get(key)
        if entry is None:
            if len(self.graphs) >= self.max_graphs:        # too many distinct patterns: run this one eagerly
                self.loss = self._eager(flags)
                return self.loss
            entry = self.graphs[key] = self._capture(flags)
        g, loss, grads, kernels = entry
        self._before()
        g.replay()
        if self.optimizer is not None and not self.dp:     # the replayed graph updated the weights through raw pointers
            self.optimizer.note_update()
        self.kernels_replayed += kernels
        for p, gr in zip(self.params, grads):              # static gradient tensors of this graph
            p.grad = gr
        self._after()
        self.loss = loss
        return loss

    def flat_grad(self):
        from .ddp import flat_grad_view
        return flat_grad_view(self.params)
