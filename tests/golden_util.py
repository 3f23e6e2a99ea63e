"""Shared helpers: load a golden case and rebuild its seeded params / inputs."""
import os

import numpy as np
import torch

from oracle import megacrn_oracle as O
from oracle.make_golden import CASES, sample_index  # noqa: F401  (pure-python tables only)

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
OUT_NAMES = ("output", "h_att", "query", "pos", "neg")


def load_case(name):
    kw, batch, t_in, batches_seen, full = CASES[name]
    d = O.Dims(**kw)
    params = O.init_params(d, seed=0)
    x, y_cov, labels = O.synthetic_batch(d, batch, t_in, seed=1234)
    gold = dict(np.load(os.path.join(GOLDEN_DIR, f"{name}.npz")))
    return d, params, (x, y_cov, labels), gold, full


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    den = float(torch.linalg.norm(b))
    return float(torch.linalg.norm(a - b)) / (den if den > 0 else 1.0)


def close_mixed(a, b, tol):
    """|a-b| <= tol * (|b| + rms(b))  -- SURVEY.md section 7.4 tolerance definition."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    rms = float(torch.sqrt(torch.mean(b * b))) if b.numel() else 0.0
    return bool(torch.all((a - b).abs() <= tol * (b.abs() + rms) + 1e-30))
