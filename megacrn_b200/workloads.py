"""The BASELINE.json configurations and the synthetic inputs of SURVEY.md section 8(d), for bench.py, smoke runs and
the stand-in trainer.  Product-side code: nothing here touches ``oracle/``.

C2 METR-LA N=207 T=12 H=64 batch 64 (the configuration the metric is quoted on), C3 PEMS-BAY N=325, C4 EXPY-TKY N=1843
T=6 batch 32, C5 N=2841 H=128 batch 256 over 8 GPUs (32 per GPU).  The class defaults M=20, d=64, cheb_k=3, L=1 apply
(model/MegaCRN.py:117-118)."""
from __future__ import annotations

from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class ModelDims:
    """Constructor arguments of ``MegaCRN`` (model/MegaCRN.py:117-118)."""
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

    def build(self, device):
        from .MegaCRN import MegaCRN
        return MegaCRN(self.num_nodes, self.input_dim, self.output_dim, self.horizon, self.rnn_units,
                       num_layers=self.num_layers, cheb_k=self.cheb_k, ycov_dim=self.ycov_dim, mem_num=self.mem_num,
                       mem_dim=self.mem_dim, cl_decay_steps=self.cl_decay_steps,
                       use_curriculum_learning=self.use_curriculum_learning).to(device)


# name -> (ModelDims kwargs, per-GPU batch, T_in, description)
CONFIGS = {
    "c1": (dict(num_nodes=207, horizon=12, rnn_units=64), 1, 12, "METR-LA N=207 T=12 H=64 batch=1"),
    "c2": (dict(num_nodes=207, horizon=12, rnn_units=64), 64, 12, "METR-LA N=207 T=12 H=64 batch=64"),
    "c3": (dict(num_nodes=325, horizon=12, rnn_units=64), 64, 12, "PEMS-BAY N=325 T=12 H=64 batch=64"),
    "c4": (dict(num_nodes=1843, horizon=6, rnn_units=64), 32, 6, "EXPY-TKY N=1843 T=6 H=64 batch=32"),
    "c5": (dict(num_nodes=2841, horizon=12, rnn_units=128), 32, 12, "synthetic N=2841 T=12 H=128 batch=256 over 8 GPUs (32 per GPU)"),
}


def config(name: str):
    kw, batch, t_in, what = CONFIGS[name]
    return ModelDims(**kw), batch, t_in, what


def synthetic_batch(d, batch: int, t_in: int, seed: int = 1234, dtype=torch.float32):
    """x ~ N(0,1) (z-scored speed), y_cov ~ U[0,1) (time of day), labels ~ N(0,1); one generator, this order
    (SURVEY.md section 8(d))."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, t_in, d.num_nodes, d.input_dim, generator=g)
    y_cov = torch.rand(batch, d.horizon, d.num_nodes, d.ycov_dim, generator=g)
    labels = torch.randn(batch, d.horizon, d.num_nodes, d.output_dim, generator=g)
    return x.to(dtype), y_cov.to(dtype), labels.to(dtype)


def forward_flops(d, batch: int, t_in: int) -> int:
    """Algorithmic forward FLOPs: identity blocks and the hoisted Chebyshev product counted once."""
    N, H, D, M, dm = d.num_nodes, d.rnn_units, d.rnn_units + d.mem_dim, d.mem_num, d.mem_dim

    def agcn(C, O):
        return 2 * 4 * N * N * batch * C + 2 * batch * N * 6 * C * O
    enc = t_in * (agcn(d.input_dim + H, 2 * H) + agcn(d.input_dim + H, H))
    cd = d.output_dim + d.ycov_dim
    dec = d.horizon * (agcn(cd + D, 2 * D) + agcn(cd + D, D))
    misc = 4 * N ** 3 + 4 * N * N * dm + 4 * N * M * dm + 2 * batch * N * H * dm + 4 * batch * N * dm * M + 2 * d.horizon * batch * N * D
    return enc + dec + misc


def agcn_flops(d, batch: int, channels: int, out: int) -> int:
    """Algorithmic FLOPs of ONE AGCN call: 2*4*N^2*B*C + 2*B*N*6C*O."""
    N = d.num_nodes
    return 2 * 4 * N * N * batch * channels + 2 * batch * N * 6 * channels * out
