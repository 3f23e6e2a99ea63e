"""Fused backward (csrc/agcn_bwd_fused.cuh) vs the per-stage backward of the same library: per-parameter rel-L2 of the
gradients on identical forwards.   usage: python tools/bwd_check.py   (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from megacrn_b200 import MegaCRN, _abi
from oracle import megacrn_oracle as O

lib = _abi.load()
dev = torch.device("cuda:0")


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def run(d, B, t_in, bwd_fused, flags):
    lib.mcrn_set_bwd_fused(bwd_fused)
    p = O.init_params(d, seed=0)
    x, y_cov, labels = O.synthetic_batch(d, B, t_in, seed=3)
    m = MegaCRN(d.num_nodes, d.input_dim, d.output_dim, d.horizon, d.rnn_units, mem_num=d.mem_num, mem_dim=d.mem_dim).to(dev)
    m.load_state_dict(p)
    m.train()
    outs = m(x.to(dev), y_cov.to(dev), labels.to(dev), teacher_forcing=flags)
    g = torch.Generator().manual_seed(5)
    ups = [torch.randn(outs[0].shape, generator=g).to(dev), torch.randn(outs[2].shape, generator=g).to(dev)]
    torch.autograd.backward([outs[0], outs[2]], ups)
    torch.cuda.synchronize()
    return [o.detach().clone() for o in outs], {k: v.grad.detach().clone() for k, v in m.named_parameters()}


cases = [("N=100 H=64 T=1 B=2", O.Dims(num_nodes=100, horizon=1, rnn_units=64), 2, 1),
         ("N=207 H=64 T=2 B=4", O.Dims(num_nodes=207, horizon=2, rnn_units=64), 4, 2),
         ("N=207 H=64 T=12 B=8", O.Dims(num_nodes=207, horizon=12, rnn_units=64), 8, 12),
         ("N=300 H=64 T=3 B=3", O.Dims(num_nodes=300, horizon=3, rnn_units=64), 3, 3),
         ("N=130 H=32 d=32 T=2 (dec only)", O.Dims(num_nodes=130, horizon=2, rnn_units=32, mem_dim=32), 2, 2)]
for title, d, B, t_in in cases:
    for flags in ([True] * d.horizon, [t % 2 == 1 for t in range(d.horizon)]):
        o0, g0 = run(d, B, t_in, 0, flags)
        for mode in (1, 2):
          o1, g1 = run(d, B, t_in, mode, flags)
          errs = sorted(((rel(g1[k], g0[k]), k) for k in g0), reverse=True)
          print(f"[{title}] mode={mode} tf={''.join('1' if f else '0' for f in flags)} out={rel(o1[0], o0[0]):.1e}  " +
                "  ".join(f"{k.split('.')[-3][:3] if k.count('.') > 1 else ''}.{k.split('.')[-2]}.{k.split('.')[-1]}={e:.1e}" for e, k in errs[:5]), flush=True)
lib.mcrn_set_bwd_fused(2)
