"""Per-parameter gradient error of the CUDA path against the CPU oracle / the reference goldens (what tests/test_gpu_parity.py
bounds): prints rel-L2 per parameter for every golden case, full C2 and C3 (B=8), so that the test tolerances can be pinned
at ~2x the measured values.   usage (GPU box): python tools/grad_errors.py > profiles/r2_grad_errors.txt"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from golden_util import CASES, load_case, rel_l2, sample_index
from oracle import megacrn_oracle as O
from megacrn_b200 import MegaCRN

dev = torch.device("cuda:0")


def model_of(d, p):
    m = MegaCRN(d.num_nodes, d.input_dim, d.output_dim, d.horizon, d.rnn_units, cheb_k=d.cheb_k, ycov_dim=d.ycov_dim,
                mem_num=d.mem_num, mem_dim=d.mem_dim).to(dev)
    m.load_state_dict(p)
    return m.train()


def upstream(o_, q_, pos, neg, labels):
    o = torch.as_tensor(o_).clone().requires_grad_(True); q = torch.as_tensor(q_).clone().requires_grad_(True)
    O.trainer_loss((o, None, q, torch.as_tensor(pos), torch.as_tensor(neg)), labels).backward()
    return o.grad, q.grad


rows = {}
for name in [n for n in CASES if n != "layers2"]:
    d, p, (x, y_cov, labels), gold, full = load_case(name)
    m = model_of(d, p)
    flags = [bool(f) for f in gold["train_flags"]]
    outs = m(x.to(dev), y_cov.to(dev), labels.to(dev), teacher_forcing=flags)
    d_out, d_q = upstream(gold["train_output"], gold["train_query"], gold["train_pos"], gold["train_neg"], labels)
    torch.autograd.backward([outs[0], outs[2]], [d_out.to(dev), d_q.to(dev)])
    e = {"fwd": rel_l2(outs[0].detach().cpu(), gold["train_output"])}
    for pn, prm in m.named_parameters():
        g = prm.grad.detach().cpu()
        if full:
            e[pn] = rel_l2(g, gold["grad_" + pn])
        else:
            flat = g.reshape(-1).numpy()
            e[pn] = rel_l2(flat[sample_index(flat.size)], gold["gsample_" + pn])
    rows["golden:" + name] = e
for tag, N, B in (("oracle:c2 B=64", 207, 64), ("oracle:c3 B=8", 325, 8)):
    d = O.Dims(num_nodes=N)
    p = O.init_params(d, seed=0)
    x, y_cov, labels = O.synthetic_batch(d, B, 12, seed=1234)
    flags = [True] * 6 + [False] * 6
    _, ref_outs, ref_grads = O.loss_and_grads(d, p, x, y_cov, labels, flags)
    m = model_of(d, p)
    outs = m(x.to(dev), y_cov.to(dev), labels.to(dev), teacher_forcing=flags)
    d_out, d_q = upstream(ref_outs[0], ref_outs[2], ref_outs[3], ref_outs[4], labels)
    torch.autograd.backward([outs[0], outs[2]], [d_out.to(dev), d_q.to(dev)])
    e = {"fwd": rel_l2(outs[0].detach().cpu(), ref_outs[0])}
    for pn, prm in m.named_parameters():
        e[pn] = rel_l2(prm.grad.cpu(), ref_grads[pn])
    rows[tag] = e
names = list(next(iter(rows.values())).keys())
print("rel-L2 error of the CUDA path (default engine: fp16 operands, hi-only forward weights)")
print(f"{'parameter':46s}" + "".join(f"{k[:16]:>18s}" for k in rows) + f"{'max':>12s}")
for n in names:
    vals = [rows[k][n] for k in rows]
    print(f"{n:46s}" + "".join(f"{v:18.2e}" for v in vals) + f"{max(vals):12.2e}")
