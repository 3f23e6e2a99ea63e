"""Is the saved forward state or the backward at fault?  Switch the engine mask between forward and backward."""
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

def run(fmask, bmask, sync=False):
    lib.mcrn_set_debug_mask(fmask)
    m = MegaCRN(d.num_nodes, d.input_dim, d.output_dim, d.horizon, d.rnn_units, mem_num=d.mem_num, mem_dim=d.mem_dim).to(dev)
    m.load_state_dict(p); m.train()
    outs = m(x.to(dev), y_cov.to(dev), labels.to(dev), teacher_forcing=flags)
    loss = O.trainer_loss(outs, labels.to(dev))
    if sync: torch.cuda.synchronize()
    lib.mcrn_set_debug_mask(bmask)
    loss.backward()
    torch.cuda.synchronize()
    errs = {}
    for pname, prm in m.named_parameters():
        flat = prm.grad.detach().cpu().reshape(-1).numpy()
        ref = gold["grad_" + pname].reshape(-1) if full else gold["gsample_" + pname]
        got = flat if full else flat[sample_index(flat.size)]
        errs[pname] = rel_l2(got, ref)
    return errs

def show(tag, e):
    print(f"{tag:28s} " + " ".join(f"{k.split('.')[-2][:3]}.{k.split('.')[-1][:4]} {v:.1e}" for k, v in e.items()))

show("fwd TC   bwd TC", run(0, 0))
show("fwd TC   bwd TC (again)", run(0, 0))
show("fwd TC   bwd TC (sync)", run(0, 0, True))
show("fwd TC   bwd SIMT", run(0, 0x7f))
show("fwd SIMT bwd TC", run(0x7f, 0))
show("fwd SIMT bwd SIMT", run(0x7f, 0x7f))
show("fwd cheb-simt, bwd TC", run(0x40, 0))
show("fwd TC, bwd dxp-simt", run(0, 0x04))
show("fwd TC, bwd dw-simt", run(0, 0x08))
show("fwd TC, bwd propT-simt", run(0, 0x10))
show("fwd TC, bwd ds-simt", run(0, 0x20))
show("fwd TC, bwd dxp+propT simt", run(0, 0x14))
lib.mcrn_set_debug_mask(0)
