"""CPU oracle for the MegaCRN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional (no nn.Module) restatement of the reference algorithm
``model/MegaCRN.py`` in plain PyTorch CPU ops.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file; the product package
``megacrn_b200`` never does (it fails loudly without its CUDA library).

Parity pin: the reference ships no tests or golden vectors, so
this oracle is pinned against outputs of the reference module itself, executed
in the build container by ``oracle/make_golden.py`` (which imports
``/root/reference/model/MegaCRN.py`` read-only) and committed under
``tests/golden/``.  ``tests/test_oracle_golden.py`` re-checks the oracle against
those fixtures wherever the tests run.

Every function cites the reference lines it restates (paths relative to the
reference checkout).  The structure is deliberately the *un-optimised* one of
the reference (identity supports, six weight blocks, per-call Chebyshev
recursion) so that the B200 kernels' algebraic shortcuts (hoisting, identity
folding, node-major layout) are checked against the original formulation.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Tensor = torch.Tensor


@dataclass(frozen=True)
class Dims:
    """Constructor arguments of the reference model (model/MegaCRN.py:117-118)."""
    num_nodes: int
    input_dim: int = 1
    output_dim: int = 1
    horizon: int = 12
    rnn_units: int = 64
    num_layers: int = 1
    cheb_k: int = 3
    ycov_dim: int = 1
    mem_num: int = 20
    mem_dim: int = 64
    cl_decay_steps: int = 2000
    use_curriculum_learning: bool = True

    @property
    def decoder_dim(self) -> int:          # model/MegaCRN.py:139
        return self.rnn_units + self.mem_dim


def param_shapes(d: Dims) -> Dict[str, Tuple[int, ...]]:
    """The 14 (for num_layers=1) state_dict entries, in registration order.

    model/MegaCRN.py:8-14 (AGCN), :149-157 (memory), :136-144 (enc/dec/proj).
    """
    shapes: Dict[str, Tuple[int, ...]] = {
        "memory.Memory": (d.mem_num, d.mem_dim),
        "memory.Wq": (d.rnn_units, d.mem_dim),
        "memory.We1": (d.num_nodes, d.mem_num),
        "memory.We2": (d.num_nodes, d.mem_num),
    }
    for i in range(d.num_layers):
        cin = d.input_dim if i == 0 else d.rnn_units
        c = cin + d.rnn_units
        p = f"encoder.dcrnn_cells.{i}."
        shapes[p + "gate.weights"] = (2 * d.cheb_k * c, 2 * d.rnn_units)
        shapes[p + "gate.bias"] = (2 * d.rnn_units,)
        shapes[p + "update.weights"] = (2 * d.cheb_k * c, d.rnn_units)
        shapes[p + "update.bias"] = (d.rnn_units,)
    dd = d.decoder_dim
    for i in range(d.num_layers):
        cin = (d.output_dim + d.ycov_dim) if i == 0 else dd
        c = cin + dd
        p = f"decoder.dcrnn_cells.{i}."
        shapes[p + "gate.weights"] = (2 * d.cheb_k * c, 2 * dd)
        shapes[p + "gate.bias"] = (2 * dd,)
        shapes[p + "update.weights"] = (2 * d.cheb_k * c, dd)
        shapes[p + "update.bias"] = (dd,)
    shapes["proj.0.weight"] = (d.output_dim, dd)
    shapes["proj.0.bias"] = (d.output_dim,)
    return shapes


def init_params(d: Dims, seed: int = 0, dtype=torch.float32) -> Dict[str, Tensor]:
    """Deterministic stand-in initialisation with the reference's distributions
    (xavier-normal matrices, zero AGCN bias; model/MegaCRN.py:13-14, :155-156).
    NOT bit-identical to constructing the reference module (different RNG
    consumption order); parity tests share an explicit state_dict instead."""
    g = torch.Generator().manual_seed(seed)
    out: Dict[str, Tensor] = {}
    for name, shp in param_shapes(d).items():
        if name.endswith("bias") and "proj" not in name:
            out[name] = torch.zeros(shp, dtype=dtype)
        elif len(shp) == 1:
            bound = 1.0 / math.sqrt(d.decoder_dim)
            out[name] = ((torch.rand(shp, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)
        else:
            fan_out, fan_in = (shp[0], shp[1])
            std = math.sqrt(2.0 / (fan_in + fan_out))
            out[name] = (torch.randn(shp, generator=g, dtype=torch.float64) * std).to(dtype)
    return out


# --------------------------------------------------------------------------
# L0: graph convolution                                   model/MegaCRN.py:16-28
# --------------------------------------------------------------------------
def chebyshev_support_set(supports: Sequence[Tensor], cheb_k: int) -> List[Tensor]:
    """[I, S, 2 S T_{k-1} - T_{k-2}, ...] for each support (model/MegaCRN.py:19-23)."""
    out: List[Tensor] = []
    for s in supports:
        ks = [torch.eye(s.shape[0], dtype=s.dtype, device=s.device), s]
        for _ in range(2, cheb_k):
            ks.append(torch.matmul(2 * s, ks[-1]) - ks[-2])
        out.extend(ks)
    return out


def agcn(x: Tensor, supports: Sequence[Tensor], weights: Tensor, bias: Tensor, cheb_k: int) -> Tensor:
    """x [B,N,C] -> [B,N,O] (model/MegaCRN.py:16-28)."""
    x_g = [torch.einsum("nm,bmc->bnc", s, x) for s in chebyshev_support_set(supports, cheb_k)]  # :24-25
    x_g = torch.cat(x_g, dim=-1)                                                                  # :26
    return torch.einsum("bni,io->bno", x_g, weights) + bias                                      # :27


# --------------------------------------------------------------------------
# L1: recurrent cell                                      model/MegaCRN.py:38-48
# --------------------------------------------------------------------------
def agcrn_cell(x: Tensor, state: Tensor, supports: Sequence[Tensor], p: Dict[str, Tensor],
               prefix: str, cheb_k: int) -> Tensor:
    hidden = state.shape[-1]
    xs = torch.cat((x, state), dim=-1)                                                      # :42
    z_r = torch.sigmoid(agcn(xs, supports, p[prefix + "gate.weights"], p[prefix + "gate.bias"], cheb_k))  # :43
    z, r = torch.split(z_r, hidden, dim=-1)                                                 # :44  (z first, r second)
    cand = torch.cat((x, z * state), dim=-1)                                                # :45
    hc = torch.tanh(agcn(cand, supports, p[prefix + "update.weights"], p[prefix + "update.bias"], cheb_k))  # :46
    return r * state + (1 - r) * hc                                                         # :47


# --------------------------------------------------------------------------
# L3: supports prologue, memory query                     model/MegaCRN.py:159-173
# --------------------------------------------------------------------------
def meta_graph_supports(p: Dict[str, Tensor]) -> List[Tensor]:
    e1 = torch.matmul(p["memory.We1"], p["memory.Memory"])                                  # :169
    e2 = torch.matmul(p["memory.We2"], p["memory.Memory"])                                  # :170
    g1 = torch.softmax(torch.relu(torch.mm(e1, e2.T)), dim=-1)                              # :171
    g2 = torch.softmax(torch.relu(torch.mm(e2, e1.T)), dim=-1)                              # :172
    return [g1, g2]


def query_memory(h_t: Tensor, p: Dict[str, Tensor]):
    mem = p["memory.Memory"]
    query = torch.matmul(h_t, p["memory.Wq"])                                               # :160
    att = torch.softmax(torch.matmul(query, mem.t()), dim=-1)                               # :161
    value = torch.matmul(att, mem)                                                          # :162
    _, ind = torch.topk(att, k=2, dim=-1)                                                   # :163
    pos = mem[ind[:, :, 0]]                                                                 # :164
    neg = mem[ind[:, :, 1]]                                                                 # :165
    return value, query, pos, neg, att, ind


def sampling_threshold(d: Dims, batches_seen) -> float:
    """model/MegaCRN.py:146-147."""
    return d.cl_decay_steps / (d.cl_decay_steps + np.exp(batches_seen / d.cl_decay_steps))


def draw_teacher_forcing(d: Dims, training: bool, batches_seen) -> List[bool]:
    """The host-side coin flips of model/MegaCRN.py:188-191, drawn from the global
    NumPy stream in the reference's order (one draw per horizon step, only in
    train mode with curriculum learning)."""
    if not (training and d.use_curriculum_learning):
        return [False] * d.horizon
    thr = sampling_threshold(d, batches_seen)
    return [bool(np.random.uniform(0, 1) < thr) for _ in range(d.horizon)]


# --------------------------------------------------------------------------
# whole forward                                           model/MegaCRN.py:168-194
# --------------------------------------------------------------------------
def forward(d: Dims, p: Dict[str, Tensor], x: Tensor, y_cov: Tensor, labels: Optional[Tensor] = None,
            teacher_forcing: Optional[Sequence[bool]] = None, return_aux: bool = False):
    """x [B,T,N,Cin], y_cov [B,T',N,ycov], labels [B,T',N,Cout] ->
    (output [B,T',N,Cout], h_att, query, pos, neg [B,N,d]).

    ``teacher_forcing[t]`` True means the decoder input of step t+1 is
    ``labels[:, t]`` (the outcome of the reference's coin flip at :189-191).
    """
    if teacher_forcing is None:
        teacher_forcing = [False] * d.horizon
    bsz = x.shape[0]
    supports = meta_graph_supports(p)                                                       # :169-173
    cur = x
    last_states = []
    for i in range(d.num_layers):                                                           # :71
        state = torch.zeros(bsz, d.num_nodes, d.rnn_units, dtype=x.dtype, device=x.device)                # :50-51, :174
        inner = []
        for t in range(cur.shape[1]):                                                       # :74
            state = agcrn_cell(cur[:, t], state, supports, p, f"encoder.dcrnn_cells.{i}.", d.cheb_k)
            inner.append(state)
        last_states.append(state)
        cur = torch.stack(inner, dim=1)                                                     # :78
    h_t = cur[:, -1]                                                                        # :176
    h_att, query, pos, neg, att, ind = query_memory(h_t, p)                                 # :178
    h_t = torch.cat([h_t, h_att], dim=-1)                                                   # :179
    ht_list = [h_t] * d.num_layers                                                          # :181
    go = torch.zeros(bsz, d.num_nodes, d.output_dim, dtype=x.dtype, device=x.device)                     # :182
    out = []
    for t in range(d.horizon):                                                              # :184
        cur_in = torch.cat([go, y_cov[:, t]], dim=-1)                                       # :185
        new_states = []
        for i in range(d.num_layers):                                                       # :109-112
            s = agcrn_cell(cur_in, ht_list[i], supports, p, f"decoder.dcrnn_cells.{i}.", d.cheb_k)
            new_states.append(s)
            cur_in = s
        ht_list = new_states
        go = torch.matmul(cur_in, p["proj.0.weight"].t()) + p["proj.0.bias"]                # :186
        out.append(go)                                                                      # :187
        if teacher_forcing[t]:                                                              # :188-191
            go = labels[:, t]
    output = torch.stack(out, dim=1)                                                        # :192
    if return_aux:
        return (output, h_att, query, pos, neg), {"att": att, "ind": ind, "supports": supports}
    return output, h_att, query, pos, neg


# --------------------------------------------------------------------------
# the trainer's loss (boundary caller)       model/traintest_MegaCRN.py:118-125
# --------------------------------------------------------------------------
def masked_mae_loss(y_pred: Tensor, y_true: Tensor) -> Tensor:
    """model/utils.py:126-133."""
    mask = (y_true != 0).to(y_pred.dtype)
    mask = mask / mask.mean()
    loss = torch.abs(y_pred - y_true) * mask
    loss = torch.where(loss != loss, torch.zeros_like(loss), loss)
    return loss.mean()


def trainer_loss(outs, labels: Tensor, scaler_mean: float = 54.0, scaler_std: float = 20.0,
                 lamb: float = 0.01, lamb1: float = 0.01) -> Tensor:
    """loss1 + lamb*triplet + lamb1*mse with pos/neg detached
    (model/traintest_MegaCRN.py:118-125, StandardScaler model/utils.py:45-54)."""
    output, _h_att, query, pos, neg = outs
    y_pred = output * scaler_std + scaler_mean
    y_true = labels * scaler_std + scaler_mean
    l1 = masked_mae_loss(y_pred, y_true)
    l2 = torch.nn.functional.triplet_margin_loss(query, pos.detach(), neg.detach(), margin=1.0)
    l3 = torch.nn.functional.mse_loss(query, pos.detach())
    return l1 + lamb * l2 + lamb1 * l3


def loss_and_grads(d: Dims, p: Dict[str, Tensor], x, y_cov, labels, teacher_forcing, **loss_kw):
    """Forward + trainer loss + autograd backward; returns (loss, outs, grads dict)."""
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    outs = forward(d, q, x, y_cov, labels, teacher_forcing)
    loss = trainer_loss(outs, labels, **loss_kw)
    grads = torch.autograd.grad(loss, list(q.values()), allow_unused=True)
    g = {k: (torch.zeros_like(v) if gi is None else gi) for (k, v), gi in zip(q.items(), grads)}
    return loss.detach(), tuple(o.detach() for o in outs), g


def synthetic_batch(d: Dims, batch: int, t_in: int, seed: int = 1234, dtype=torch.float32):
    """SURVEY.md section 8(d) synthetic inputs: x ~ N(0,1), y_cov ~ U[0,1), labels ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, t_in, d.num_nodes, d.input_dim, generator=g)
    y_cov = torch.rand(batch, d.horizon, d.num_nodes, d.ycov_dim, generator=g)
    labels = torch.randn(batch, d.horizon, d.num_nodes, d.output_dim, generator=g)
    return x.to(dtype), y_cov.to(dtype), labels.to(dtype)


# ---- host input pipeline of the reference (model/utils.py:6-43, model/traintest_MegaCRN.py:33-48), restated ----
class DataLoaderOracle:
    """numpy restatement of the reference ``DataLoader``: pad with the last sample (:17-22), one ``np.random.permutation``
    when shuffling (:25-27), ``num_batch = size // batch_size`` full batches (:24, :33-40)."""

    def __init__(self, xs, ys, batch_size, pad_with_last_sample=True, shuffle=False):
        import numpy as np
        self.batch_size = batch_size
        if pad_with_last_sample:
            num_padding = (batch_size - (len(xs) % batch_size)) % batch_size
            xs = np.concatenate([xs, np.repeat(xs[-1:], num_padding, axis=0)], axis=0)
            ys = np.concatenate([ys, np.repeat(ys[-1:], num_padding, axis=0)], axis=0)
        self.size = len(xs)
        self.num_batch = int(self.size // self.batch_size)
        if shuffle:
            permutation = np.random.permutation(self.size)
            xs, ys = xs[permutation], ys[permutation]
        self.xs, self.ys = xs, ys

    def batches(self, input_dim=1, output_dim=1):
        """(x, y, y_cov) float32 arrays per batch, as prepare_x_y slices them (traintest:41-47)."""
        import numpy as np
        for i in range(self.num_batch):
            s, e = self.batch_size * i, min(self.size, self.batch_size * (i + 1))
            x, y = self.xs[s:e], self.ys[s:e]
            yield (x[..., :input_dim].astype(np.float32), y[..., :output_dim].astype(np.float32),
                   y[..., output_dim:].astype(np.float32))
