"""Drop-in replacement for the reference's ``model/MegaCRN.py``.

``from MegaCRN import MegaCRN`` (model/traintest_MegaCRN.py:15) resolves to this file when
``megacrn_b200`` is first on ``sys.path`` (see megacrn_b200/launch_traintest.py).  Same
constructor (model/MegaCRN.py:117-118), same ``forward(x, y_cov, labels, batches_seen)``
-> ``(output, h_att, query, pos, neg)`` (:168, :194), same ``state_dict`` keys and shapes (14 for the default
``num_layers=1``, 8 more per stacked layer),
same consumption of ``np.random.uniform`` for scheduled sampling (:188-191).

All arithmetic runs in hand-written sm_100a kernels behind the C ABI of
``libmegacrn_b200.so``; this file is plumbing (parameter containers, one
``torch.autograd.Function``, workspace management).  There is NO CPU path: tensors must
live on a CUDA device and the shared library must be built.
"""
from __future__ import annotations

import weakref

import numpy as np
import torch
import torch.nn as nn

from . import _abi


# ---- parameter containers with the reference's module/parameter names -----------------
class AGCN(nn.Module):
    """Parameters of model/MegaCRN.py:8-14 (weights [2*cheb_k*dim_in, dim_out], bias)."""

    def __init__(self, dim_in, dim_out, cheb_k):
        super().__init__()
        self.cheb_k = cheb_k
        self.weights = nn.Parameter(torch.empty(2 * cheb_k * dim_in, dim_out))
        self.bias = nn.Parameter(torch.empty(dim_out))
        nn.init.xavier_normal_(self.weights)
        nn.init.constant_(self.bias, val=0)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("megacrn_b200.AGCN is a parameter container; the fused kernels run from MegaCRN.forward")


class AGCRNCell(nn.Module):
    """model/MegaCRN.py:30-36."""

    def __init__(self, node_num, dim_in, dim_out, cheb_k):
        super().__init__()
        self.node_num = node_num
        self.hidden_dim = dim_out
        self.gate = AGCN(dim_in + dim_out, 2 * dim_out, cheb_k)
        self.update = AGCN(dim_in + dim_out, dim_out, cheb_k)

    def init_hidden_state(self, batch_size):
        return torch.zeros(batch_size, self.node_num, self.hidden_dim)


class _CellStack(nn.Module):
    def __init__(self, node_num, dim_in, dim_out, cheb_k, num_layers, what):
        super().__init__()
        assert num_layers >= 1, f"At least one DCRNN layer in the {what}."
        self.node_num = node_num
        self.input_dim = dim_in
        self.num_layers = num_layers
        self.dcrnn_cells = nn.ModuleList([AGCRNCell(node_num, dim_in, dim_out, cheb_k)])
        for _ in range(1, num_layers):
            self.dcrnn_cells.append(AGCRNCell(node_num, dim_out, dim_out, cheb_k))


class ADCRNN_Encoder(_CellStack):
    """model/MegaCRN.py:53-63 (parameters only)."""

    def __init__(self, node_num, dim_in, dim_out, cheb_k, num_layers):
        super().__init__(node_num, dim_in, dim_out, cheb_k, num_layers, "Encoder")


class ADCRNN_Decoder(_CellStack):
    """model/MegaCRN.py:91-101 (parameters only)."""

    def __init__(self, node_num, dim_in, dim_out, cheb_k, num_layers):
        super().__init__(node_num, dim_in, dim_out, cheb_k, num_layers, "Decoder")


class _Workspace:
    """One device workspace; ``busy`` while a saved forward awaits its backward."""

    def __init__(self, nbytes, device):
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.nbytes = nbytes
        self.busy = False
        self.prologue_key = None      # validity key of the parameter-only prologue held by this workspace (eval fast path)


def _release(ws):
    ws.busy = False


class _MegaCRNFunction(torch.autograd.Function):
    """forward = mcrn_forward, backward = mcrn_backward (include/megacrn_b200.h)."""

    @staticmethod
    def forward(ctx, module, need_grad, x, y_cov, labels, tf, *params):
        lib = _abi.load()
        dims = module._dims(x)
        flags = _abi.MCRN_FWD_SAVE_FOR_BACKWARD if need_grad else 0
        ws = module._workspace(dims, flags, x.device)
        # eval fast path: supports, folded weights and their operand copies depend on the parameters only; they
        # are reused while no parameter has been modified in place or re-allocated and the library mode is unchanged
        # (`_param_epoch` counts the updates that go through raw pointers -- FusedClipAdam / a replayed CUDA graph -- and
        # therefore bump no tensor version)
        key = (tuple(getattr(dims, f) for f, _ in dims._fields_), flags, lib.mcrn_mode_epoch(), module._param_epoch,
               tuple((p.data_ptr(), p._version) for p in params))
        if not need_grad and not module.training and ws.prologue_key == key:
            flags |= _abi.MCRN_FWD_REUSE_PROLOGUE
        ws.prologue_key = None
        B, N, d = x.shape[0], module.num_nodes, module.mem_dim
        output = torch.empty(B, module.horizon, N, module.output_dim, device=x.device, dtype=torch.float32)
        h_att, query, pos, neg = (torch.empty(B, N, d, device=x.device, dtype=torch.float32) for _ in range(4))
        prm = _abi.make_params(params[:14])
        upper = _abi.make_layer_params(params[14:])          # stacked cells of layers >= 1 (None for num_layers == 1)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            if upper is None:
                st = lib.mcrn_forward(dims, prm, x.data_ptr(), y_cov.data_ptr(), _abi.ptr(labels), tf,
                                      output.data_ptr(), h_att.data_ptr(), query.data_ptr(), pos.data_ptr(),
                                      neg.data_ptr(), ws.buf.data_ptr(), ws.nbytes, flags, stream)
            else:
                st = lib.mcrn_forward_layers(dims, prm, _abi.layer_ptr(upper), x.data_ptr(), y_cov.data_ptr(), _abi.ptr(labels),
                                             tf, output.data_ptr(), h_att.data_ptr(), query.data_ptr(), pos.data_ptr(),
                                             neg.data_ptr(), ws.buf.data_ptr(), ws.nbytes,
                                             flags & ~_abi.MCRN_FWD_REUSE_PROLOGUE, stream)
        _abi.check(st, "mcrn_forward")
        ws.prologue_key = key
        if need_grad:
            ws.busy = True
            weakref.finalize(ctx, _release, ws)
            ctx.ws, ctx.dims, ctx.tf = ws, dims, tf
            ctx.inputs = (x, y_cov, labels)
            ctx.params = params
            ctx.param_versions = tuple(p._version for p in params)
            ctx.done = False
        ctx.set_materialize_grads(False)
        return output, h_att, query, pos, neg

    @staticmethod
    def backward(ctx, d_out, d_hatt, d_query, d_pos, d_neg):
        lib = _abi.load()
        params = ctx.params
        if ctx.done:
            raise RuntimeError("megacrn_b200: backward through this forward a second time -- the saved activations live in "
                               "a workspace that is released (and its accumulators consumed) by the first backward; "
                               "run the forward again (retain_graph is not supported)")
        if tuple(p._version for p in params) != ctx.param_versions:
            raise RuntimeError("megacrn_b200: a parameter was modified in place between forward and backward")
        x, y_cov, labels = ctx.inputs
        # one flat fp32 buffer aliased by all the gradients -> a single NCCL all-reduce per step (ddp.py)
        sizes = [p.numel() for p in params]
        padded = [(s + 63) // 64 * 64 for s in sizes]
        flat = torch.empty(sum(padded), device=x.device, dtype=torch.float32)
        grads, off = [], 0
        for p, s, ps in zip(params, sizes, padded):
            grads.append(flat[off:off + s].view(p.shape))
            off += ps
        cont = lambda t: None if t is None else t.contiguous()
        d_out, d_hatt, d_query, d_pos, d_neg = map(cont, (d_out, d_hatt, d_query, d_pos, d_neg))
        stream = torch.cuda.current_stream(x.device).cuda_stream
        upper, upper_grads = _abi.make_layer_params(params[14:]), _abi.make_layer_params(grads[14:])
        with torch.cuda.device(x.device):
            if upper is None:
                st = lib.mcrn_backward(ctx.dims, _abi.make_params(params), x.data_ptr(), y_cov.data_ptr(),
                                       _abi.ptr(labels), ctx.tf, _abi.ptr(d_out), _abi.ptr(d_hatt), _abi.ptr(d_query),
                                       _abi.ptr(d_pos), _abi.ptr(d_neg), _abi.make_params(grads), ctx.ws.buf.data_ptr(),
                                       ctx.ws.nbytes, stream)
            else:
                st = lib.mcrn_backward_layers(ctx.dims, _abi.make_params(params[:14]), _abi.layer_ptr(upper), x.data_ptr(),
                                              y_cov.data_ptr(), _abi.ptr(labels), ctx.tf, _abi.ptr(d_out), _abi.ptr(d_hatt),
                                              _abi.ptr(d_query), _abi.ptr(d_pos), _abi.ptr(d_neg),
                                              _abi.make_params(grads[:14]), _abi.layer_ptr(upper_grads),
                                              ctx.ws.buf.data_ptr(), ctx.ws.nbytes, stream)
        _abi.check(st, "mcrn_backward")
        ctx.done = True
        ctx.ws.busy = False
        return (None, None, None, None, None, None) + tuple(grads)


class MegaCRN(nn.Module):
    """Same interface as the reference class (model/MegaCRN.py:116-194)."""

    def __init__(self, num_nodes, input_dim, output_dim, horizon, rnn_units, num_layers=1, cheb_k=3,
                 ycov_dim=1, mem_num=20, mem_dim=64, cl_decay_steps=2000, use_curriculum_learning=True):
        super().__init__()
        self.num_nodes = num_nodes
        self.input_dim = input_dim
        self.rnn_units = rnn_units
        self.output_dim = output_dim
        self.horizon = horizon
        self.num_layers = num_layers
        self.cheb_k = cheb_k
        self.ycov_dim = ycov_dim
        self.cl_decay_steps = cl_decay_steps
        self.use_curriculum_learning = use_curriculum_learning
        # construction order mirrors the reference (:130-144) so that the same torch seed
        # yields the same initial weights
        self.mem_num = mem_num
        self.mem_dim = mem_dim
        self.memory = self.construct_memory()
        self.encoder = ADCRNN_Encoder(num_nodes, input_dim, rnn_units, cheb_k, num_layers)
        self.decoder_dim = rnn_units + mem_dim
        self.decoder = ADCRNN_Decoder(num_nodes, output_dim + ycov_dim, self.decoder_dim, cheb_k, num_layers)
        self.proj = nn.Sequential(nn.Linear(self.decoder_dim, output_dim, bias=True))
        self._ws_pool = {}
        self._param_epoch = 0            # bumped by every parameter update that bypasses torch's version counters
        self.last_teacher_forcing = None
        if not 1 <= num_layers <= _abi.MAX_LAYERS:
            raise NotImplementedError(f"megacrn_b200 implements num_layers 1..{_abi.MAX_LAYERS} (got {num_layers})")

    def compute_sampling_threshold(self, batches_seen):
        """model/MegaCRN.py:146-147."""
        return self.cl_decay_steps / (self.cl_decay_steps + np.exp(batches_seen / self.cl_decay_steps))

    def construct_memory(self):
        """model/MegaCRN.py:149-157."""
        memory_dict = nn.ParameterDict()
        memory_dict["Memory"] = nn.Parameter(torch.randn(self.mem_num, self.mem_dim), requires_grad=True)
        memory_dict["Wq"] = nn.Parameter(torch.randn(self.rnn_units, self.mem_dim), requires_grad=True)
        memory_dict["We1"] = nn.Parameter(torch.randn(self.num_nodes, self.mem_num), requires_grad=True)
        memory_dict["We2"] = nn.Parameter(torch.randn(self.num_nodes, self.mem_num), requires_grad=True)
        for param in memory_dict.values():
            nn.init.xavier_normal_(param)
        return memory_dict

    # ---- plumbing -----------------------------------------------------------------------
    def note_parameter_update(self):
        """Called by every updater that writes the parameters through raw pointers (FusedClipAdam.step, a replayed
        CUDA graph that contains it): invalidates the eval fast path's cached prologue."""
        self._param_epoch += 1

    def train(self, mode=True):
        # leaving / entering training drops every cached prologue: whatever happened to the weights in between, the
        # next eval forward rebuilds supports, folded weights and operand copies from the live parameters
        if mode != self.training:
            for pool in self._ws_pool.values():
                for ws in pool:
                    ws.prologue_key = None
        return super().train(mode)

    def _ordered_params(self):
        """The tensors in C-ABI order (_abi.param_keys): the 14 of ``mcrn_params``, then 8 per stacked layer."""
        def cell(c):
            return (c.gate.weights, c.gate.bias, c.update.weights, c.update.bias)
        out = (self.memory["Memory"], self.memory["Wq"], self.memory["We1"], self.memory["We2"]) + \
            cell(self.encoder.dcrnn_cells[0]) + cell(self.decoder.dcrnn_cells[0]) + (self.proj[0].weight, self.proj[0].bias)
        for i in range(1, self.num_layers):
            out += cell(self.encoder.dcrnn_cells[i]) + cell(self.decoder.dcrnn_cells[i])
        return out

    def _dims(self, x):
        return _abi.Dims(batch=x.shape[0], num_nodes=self.num_nodes, seq_len=x.shape[1], horizon=self.horizon,
                         input_dim=self.input_dim, output_dim=self.output_dim, ycov_dim=self.ycov_dim,
                         rnn_units=self.rnn_units, num_layers=self.num_layers, cheb_k=self.cheb_k,
                         mem_num=self.mem_num, mem_dim=self.mem_dim)

    def _workspace(self, dims, flags, device):
        lib = _abi.load()
        nbytes = lib.mcrn_workspace_bytes(dims, flags)
        if nbytes == 0:
            _abi.check(-1, "mcrn_workspace_bytes")
        key = (str(device), nbytes)
        pool = self._ws_pool.setdefault(key, [])
        for ws in pool:
            if not ws.busy:
                return ws
        ws = _Workspace(nbytes, device)
        pool.append(ws)
        return ws

    def draw_teacher_forcing(self, batches_seen):
        """The coin flips of model/MegaCRN.py:188-191, drawn up front in the reference's order so the
        global NumPy stream stays aligned with the reference (exactly `horizon` draws in train mode)."""
        if not (self.training and self.use_curriculum_learning):
            return None
        thr = self.compute_sampling_threshold(batches_seen)
        return [bool(np.random.uniform(0, 1) < thr) for _ in range(self.horizon)]

    def forward(self, x, y_cov, labels=None, batches_seen=None, teacher_forcing=None):
        if not x.is_cuda:
            raise RuntimeError("megacrn_b200.MegaCRN has no CPU path: move the module and its inputs to a B200 "
                               "(`.to('cuda')`); the reference implementation is the CPU path")
        assert x.dim() == 4 and x.shape[2] == self.num_nodes and x.shape[3] == self.input_dim, x.shape
        assert y_cov.shape[1] >= self.horizon and y_cov.shape[2] == self.num_nodes
        flags = teacher_forcing if teacher_forcing is not None else self.draw_teacher_forcing(batches_seen)
        self.last_teacher_forcing = flags
        if flags is not None and any(flags) and labels is None:
            raise ValueError("teacher forcing selected a label step but labels is None")
        f32 = lambda t: None if t is None else t.to(device=x.device, dtype=torch.float32).contiguous()
        x, y_cov, labels = f32(x), f32(y_cov[:, :self.horizon]), f32(labels)
        tf = _abi.tf_bytes(flags, self.horizon)
        params = self._ordered_params()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in params)
        return _MegaCRNFunction.apply(self, need_grad, x, y_cov, labels, tf, *params)


def print_params(model):
    """model/MegaCRN.py:196-205."""
    param_count = 0
    print("Trainable parameter list:")
    for name, param in model.named_parameters():
        if param.requires_grad:
            print(name, param.shape, param.numel())
            param_count += param.numel()
    print(f"In total: {param_count} trainable parameters. \n")
