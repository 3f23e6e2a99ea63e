"""Fused AGCN kernel (csrc/agcn_fused.cuh) vs the per-stage GEMM path of the same library, and vs the CPU oracle.
usage: python tools/fused_check.py            (run on the GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from megacrn_b200 import MegaCRN, _abi
from oracle import megacrn_oracle as O

lib = _abi.load()
dev = torch.device("cuda:0")


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def run(d, B, t_in, fused, parts, train):
    lib.mcrn_set_fused(fused, parts)
    p = O.init_params(d, seed=0)
    x, y_cov, labels = O.synthetic_batch(d, B, t_in, seed=3)
    m = MegaCRN(d.num_nodes, d.input_dim, d.output_dim, d.horizon, d.rnn_units, mem_num=d.mem_num, mem_dim=d.mem_dim).to(dev)
    m.load_state_dict(p)
    if train:
        m.train()
        flags = [t % 2 == 0 for t in range(d.horizon)]
        outs = m(x.to(dev), y_cov.to(dev), labels.to(dev), teacher_forcing=flags)
        g = torch.Generator().manual_seed(5)
        ups = [torch.randn(outs[0].shape, generator=g).to(dev), torch.randn(outs[2].shape, generator=g).to(dev)]
        torch.autograd.backward([outs[0], outs[2]], ups)
        torch.cuda.synchronize()
        grads = {k: v.grad.detach().clone() for k, v in m.named_parameters()}
        return [o.detach().clone() for o in outs], grads
    m.eval()
    with torch.no_grad():
        outs = m(x.to(dev), y_cov.to(dev))
    torch.cuda.synchronize()
    return [o.clone() for o in outs], None


cases = [("enc1 N=100 H=64 T=1", O.Dims(num_nodes=100, horizon=1, rnn_units=64), 2, 1),
         ("N=207 H=64 T=1", O.Dims(num_nodes=207, horizon=1, rnn_units=64), 4, 1),
         ("N=207 H=64 T=12 B=8", O.Dims(num_nodes=207, horizon=12, rnn_units=64), 8, 12),
         ("N=300 H=64 T=3 B=3", O.Dims(num_nodes=300, horizon=3, rnn_units=64), 3, 3),
         ("N=130 H=32(d=96) T=2", O.Dims(num_nodes=130, horizon=2, rnn_units=32, mem_dim=32), 2, 2)]
names = ["output", "h_att", "query", "pos", "neg"]
for title, d, B, t_in in cases:
    for train in (False, True):
        try:
            ref, gref = run(d, B, t_in, 0, 2, train)
            for fmode, parts in ((1, 2), (2, 2), (2, 1)):
                got, gg = run(d, B, t_in, fmode, parts, train)
                line = " ".join(f"{n}={rel(a, b):.2e}" for n, a, b in zip(names[:3], got, ref))
                if train:
                    worst = max((rel(gg[k], gref[k]), k) for k in gg)
                    line += f"  worst-grad={worst[0]:.2e} ({worst[1]})"
                print(f"[{title}] train={int(train)} fused={fmode} parts={parts}: vs unfused {line}", flush=True)
        except Exception as e:
            print(f"[{title}] train={int(train)} FAILED: {e}", flush=True)
            raise
lib.mcrn_set_fused(2, 2)
