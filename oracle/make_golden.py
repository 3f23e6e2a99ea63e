"""Generate golden vectors by EXECUTING THE UNMODIFIED REFERENCE in the build container.

    python oracle/make_golden.py            # writes tests/golden/*.npz

Imports ``/root/reference/model/MegaCRN.py`` read-only (nothing is copied), loads
deterministic parameters (``oracle.megacrn_oracle.init_params``) through the
reference's own ``load_state_dict``, runs the reference ``forward`` (eval mode
and train mode with the reference's own ``np.random`` coin flips), the trainer's
loss (model/traintest_MegaCRN.py:118-125, model/utils.py:126-133) and
``loss.backward()``, and stores outputs and gradients.

Parameters and inputs are NOT stored: they are regenerated from seeds by the
oracle module, so fixtures stay small.  The GPU box has no /root/reference;
tests there read only the committed .npz files.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("MEGACRN_REFERENCE", "/root/reference")

from oracle import megacrn_oracle as O  # noqa: E402

# name -> (Dims kwargs, batch, t_in, batches_seen for the train case, full_grads)
CASES = {
    "tiny":   (dict(num_nodes=13, horizon=4, rnn_units=8, mem_num=5, mem_dim=8), 3, 5, 20000, True),
    "small":  (dict(num_nodes=37, horizon=6, rnn_units=16, mem_num=7, mem_dim=12), 4, 6, 4200, True),
    "odd":    (dict(num_nodes=50, horizon=3, rnn_units=12, mem_num=6, mem_dim=20, input_dim=2, cheb_k=3), 2, 4, 0, True),
    "layers2": (dict(num_nodes=11, horizon=3, rnn_units=8, mem_num=4, mem_dim=8, num_layers=2), 2, 3, 20000, True),
    "c1":     (dict(num_nodes=207, horizon=12, rnn_units=64), 1, 12, 20000, False),   # BASELINE.json configs[0]
    "c2b4":   (dict(num_nodes=207, horizon=12, rnn_units=64), 4, 12, 0, False),       # configs[1] shape, 4 sequences
}
GRAD_SAMPLES = 256


def sample_index(numel: int, k: int = GRAD_SAMPLES) -> np.ndarray:
    """Deterministic sample positions shared by generator and tests."""
    if numel <= k:
        return np.arange(numel)
    return (np.arange(k, dtype=np.int64) * 2654435761 % numel).astype(np.int64)


def load_reference():
    sys.path.insert(0, os.path.join(REF, "model"))
    import MegaCRN as ref_mod  # the reference's own file
    return ref_mod


def reference_trainer_loss(output, query, pos, neg, labels):
    """The trainer's loss, built from torch's own modules exactly as the
    reference script does (model/traintest_MegaCRN.py:118-125)."""
    mean, std = 54.0, 20.0
    y_pred = output * std + mean
    y_true = labels * std + mean
    mask = (y_true != 0).float()
    mask /= mask.mean()
    loss = torch.abs(y_pred - y_true) * mask
    loss[loss != loss] = 0
    l1 = loss.mean()
    l2 = torch.nn.TripletMarginLoss(margin=1.0)(query, pos.detach(), neg.detach())
    l3 = torch.nn.MSELoss()(query, pos.detach())
    return l1 + 0.01 * l2 + 0.01 * l3


def run_case(ref_mod, name: str):
    kw, batch, t_in, batches_seen, full = CASES[name]
    d = O.Dims(**kw)
    params = O.init_params(d, seed=0)
    x, y_cov, labels = O.synthetic_batch(d, batch, t_in, seed=1234)
    model = ref_mod.MegaCRN(num_nodes=d.num_nodes, input_dim=d.input_dim, output_dim=d.output_dim,
                            horizon=d.horizon, rnn_units=d.rnn_units, num_layers=d.num_layers,
                            cheb_k=d.cheb_k, ycov_dim=d.ycov_dim, mem_num=d.mem_num, mem_dim=d.mem_dim,
                            cl_decay_steps=d.cl_decay_steps,
                            use_curriculum_learning=d.use_curriculum_learning)
    missing = model.load_state_dict(params, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    out = {}
    # ---- eval mode ----------------------------------------------------
    model.eval()
    with torch.no_grad():
        o = model(x, y_cov)
    for k, v in zip(("output", "h_att", "query", "pos", "neg"), o):
        out["eval_" + k] = v.numpy()
    # ---- train mode, the reference's own coin flips -------------------
    model.train()
    np.random.seed(7)
    thr = model.compute_sampling_threshold(batches_seen)
    st = np.random.get_state()
    flags = np.array([np.random.uniform(0, 1) < thr for _ in range(d.horizon)])
    np.random.set_state(st)                      # rewind: the reference now draws the same values
    model.zero_grad()
    o = model(x, y_cov, labels, batches_seen)
    loss = reference_trainer_loss(o[0], o[2], o[3], o[4], labels)
    loss.backward()
    out["train_flags"] = flags
    out["train_batches_seen"] = np.int64(batches_seen)
    out["train_loss"] = loss.detach().numpy()
    for k, v in zip(("output", "h_att", "query", "pos", "neg"), o):
        out["train_" + k] = v.detach().numpy()
    for pname, prm in model.named_parameters():
        g = prm.grad.detach().numpy().astype(np.float32)
        if full:
            out["grad_" + pname] = g
        else:
            flat = g.reshape(-1)
            out["gsample_" + pname] = flat[sample_index(flat.size)]
            out["gnorm_" + pname] = np.float64(np.linalg.norm(flat.astype(np.float64)))
            out["gsum_" + pname] = np.float64(flat.astype(np.float64).sum())
    return out


def main():
    ref_mod = load_reference()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    for name in CASES:
        res = run_case(ref_mod, name)
        path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
        np.savez_compressed(path, **res)
        print(f"{name}: {len(res)} arrays, {os.path.getsize(path)/1024:.1f} KiB, "
              f"loss={float(res['train_loss']):.6f} flags={res['train_flags'].astype(int).tolist()}")


if __name__ == "__main__":
    main()
