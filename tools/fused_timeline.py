"""Timeline of one CTA of the fused AGCN kernel (clock64 stamps recorded by csrc/agcn_fused.cuh when
mcrn_debug_fused_timeline is armed).  Runs a 1-step encoder + 1-step decoder forward at the C2 shape; the LAST fused
launch (decoder update) and, with --gate, the decoder gate are reported.   usage: python tools/fused_timeline.py [H] [N] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from megacrn_b200 import MegaCRN, _abi
from oracle import megacrn_oracle as O

H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 207
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
lib = _abi.load()
dev = torch.device("cuda:0")
d = O.Dims(num_nodes=N, horizon=1, rnn_units=H)
p = O.init_params(d, seed=0)
x, y_cov, labels = O.synthetic_batch(d, B, 1, seed=3)
m = MegaCRN(d.num_nodes, d.input_dim, d.output_dim, d.horizon, d.rnn_units, mem_num=d.mem_num, mem_dim=d.mem_dim).to(dev)
m.load_state_dict(p)
m.train()
xs, ys, ls = x.to(dev), y_cov.to(dev), labels.to(dev)
for _ in range(3):
    m(xs, ys, ls, teacher_forcing=[True])
torch.cuda.synchronize()
# the decoder gate launch is the 3rd fused launch of a forward, the decoder update the 4th: record all, keep per-launch
# copies by running the forward with the buffer armed and reading it after each variant
WHICH = int(os.environ.get("WHICH", "2"))      # 0 enc gate, 1 enc update, 2 dec gate, 3 dec update
slots = torch.zeros(512 + 2 * 1024, dtype=torch.int64, device=dev)
lib.mcrn_debug_fused_timeline(slots.data_ptr(), WHICH)
m(xs, ys, ls, teacher_forcing=[True])
torch.cuda.synchronize()
lib.mcrn_debug_fused_timeline(None, -1)
t = slots.cpu().tolist()
t0 = t[0]
rel = lambda v: (v - t0) if v else None
print(f"fused launch #{WHICH} (0 enc gate, 1 enc update, 2 dec gate, 3 dec update), H={H} N={N} B={B}, CTA (0,0); cycles since CTA start")
print("prologue done", rel(t[1]))
items = [(i, rel(t[2 + i])) for i in range(200) if t[2 + i]]
prod = [(i, rel(t[240 + i])) for i in range(200) if t[240 + i]]
print("items:", len(items))
prev = rel(t[1])
for (i, v), (_, pv) in zip(items, prod):
    print(f"  item {i:3d}: slot free (producer) {pv:7d}   operands landed (MMA) {v:7d}   (+{v - prev})")
    prev = v
for k in range(5):
    a, b = t[210 + 4 * k], t[211 + 4 * k]
    if a:
        print(f"P_{k}: full seen {rel(a)}  rounded {rel(b)}  (+{b - a})")
print("producer done", rel(t[232]), " MMA issuer done", rel(t[233]), " acc_full seen", rel(t[230]), " epilogue done", rel(t[231]))

import numpy as np
ct = np.array(t[512:512 + 2 * 2 * B]).reshape(-1, 2)
ct = ct[ct[:, 0] > 0]
if len(ct):
    s0 = ct[:, 0].min()
    st, en = ct[:, 0] - s0, ct[:, 1] - s0
    print(f"per-CTA wall clock (ns, {len(ct)} CTAs): start min/median/max {st.min()} {int(np.median(st))} {st.max()}   end min/median/max {en.min()} {int(np.median(en))} {en.max()}   duration min/median/max {(en-st).min()} {int(np.median(en-st))} {(en-st).max()}")
    odd = ct[1::2]; even = ct[0::2]
    print(f"  tile 0 CTAs duration median {int(np.median(even[:,1]-even[:,0]))}   tile 1 CTAs duration median {int(np.median(odd[:,1]-odd[:,0]))}")
