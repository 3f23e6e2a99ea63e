"""Where a forward's time goes at launch granularity: every fused AGCN launch of one eval forward (graph replay or eager)
records {earliest CTA start, latest CTA end} in %globaltimer ns (mcrn_debug_launch_spans); printed next to the CUDA-event
time of the whole forward.   usage: python tools/launch_spans.py [config] [graph|eager] [train]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from megacrn_b200 import _abi
from megacrn_b200.workloads import config, synthetic_batch

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
mode = sys.argv[2] if len(sys.argv) > 2 else "graph"
train = len(sys.argv) > 3
lib = _abi.load()
dev = torch.device("cuda:0")
d, B, t_in, _ = config(cfg)
torch.manual_seed(0)
m = d.build(dev)
m.train() if train else m.eval()
x, y_cov, labels = (t.to(dev) for t in synthetic_batch(d, B, t_in))
flags = [True] * d.horizon if train else None
NL = 400
slots = torch.zeros(2 * NL, dtype=torch.int64, device=dev)


def arm():
    slots.view(-1, 2)[:, 0] = torch.iinfo(torch.int64).max
    slots.view(-1, 2)[:, 1] = 0
    lib.mcrn_debug_launch_spans(slots.data_ptr(), NL)


step_mode = mode == "step"
if step_mode:
    from megacrn_b200.train_step import train_step
    m.train(); flags = [True] * d.horizon; mode = "eager"


def fwd():
    if step_mode:
        for q in m.parameters():
            q.grad = None
        return train_step(m, x, y_cov, labels, teacher_forcing=flags, scaler_mean=54.0, scaler_std=20.0)
    with torch.no_grad():
        return m(x, y_cov, labels if train else None, teacher_forcing=flags)

for _ in range(3):
    fwd()
torch.cuda.synchronize()
arm()
if mode == "graph":
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fwd()
    torch.cuda.current_stream().wait_stream(s)
    arm()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fwd()
    run = g.replay
else:
    run = fwd
lib.mcrn_debug_launch_spans(None, 0) if mode == "graph" else None
for _ in range(3):
    run()
torch.cuda.synchronize()
slots.view(-1, 2)[:, 0] = torch.iinfo(torch.int64).max
slots.view(-1, 2)[:, 1] = 0
if mode != "graph":
    arm()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
lib.mcrn_debug_launch_spans(None, 0)
sp = slots.cpu().numpy().reshape(-1, 2)
sp = sp[sp[:, 1] > 0]
t0 = sp[:, 0].min()
dur = sp[:, 1] - sp[:, 0]
gap = sp[1:, 0] - sp[:-1, 1]
print(f"{cfg} {mode} {'train' if train else 'eval'} forward: {e0.elapsed_time(e1) * 1e3:.1f} us by CUDA events; {len(sp)} fused launches")
print(f"  first CTA start -> last CTA end: {(sp[:, 1].max() - t0) / 1e3:.1f} us; sum of launch spans {dur.sum() / 1e3:.1f} us; sum of gaps between consecutive fused launches {gap.sum() / 1e3:.1f} us")
try:
    print("  span us per launch:", " ".join(f"{v / 1e3:.1f}" for v in dur))
    print("  gap  us after launch:", " ".join(f"{v / 1e3:.1f}" for v in gap))
except BrokenPipeError:
    pass
