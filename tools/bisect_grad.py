"""Bisect which GEMM call-site class of the tcgen05 path produces a gradient mismatch (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from golden_util import load_case, rel_l2, sample_index
from oracle import megacrn_oracle as O
from megacrn_b200 import MegaCRN, _abi

lib = _abi.load()
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "c2b4"
d, p, (x, y_cov, labels), gold, full = load_case(name)
flags = [bool(f) for f in gold["train_flags"]]
NAMES = ["propagate", "gate/update", "make_dxp", "acc_dw", "propagate_T", "acc_ds", "cheb", "no-3x"]
for mask in [0] + [1 << i for i in range(8)] + [0x7f]:
    lib.mcrn_set_debug_mask(mask)
    m = MegaCRN(d.num_nodes, d.input_dim, d.output_dim, d.horizon, d.rnn_units, mem_num=d.mem_num, mem_dim=d.mem_dim).to(dev)
    m.load_state_dict(p); m.train()
    outs = m(x.to(dev), y_cov.to(dev), labels.to(dev), teacher_forcing=flags)
    O.trainer_loss(outs, labels.to(dev)).backward()
    errs = {}
    for pname, prm in m.named_parameters():
        g = prm.grad.detach().cpu()
        if full:
            errs[pname] = rel_l2(g, gold["grad_" + pname])
        else:
            flat = g.reshape(-1).numpy()
            errs[pname] = rel_l2(flat[sample_index(flat.size)], gold["gsample_" + pname])
    label = "none" if mask == 0 else ("all" if mask == 0x7f else NAMES[mask.bit_length() - 1])
    print(f"simt[{label:12s}] out {rel_l2(outs[0].detach().cpu(), gold['train_output']):.1e} " +
          " ".join(f"{k.split('.')[-2][:3]}.{k.split('.')[-1][:4]} {v:.1e}" for k, v in errs.items()))
lib.mcrn_set_debug_mask(0)
