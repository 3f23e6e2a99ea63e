"""CPU: the oracle restatement vs golden vectors produced by the reference itself."""
import numpy as np
import pytest
import torch

from oracle import megacrn_oracle as O
from golden_util import CASES, OUT_NAMES, load_case, rel_l2, sample_index

TOL = 2e-5   # fp32 CPU restatement vs fp32 reference: op order differs only inside BLAS


@pytest.mark.parametrize("name", list(CASES))
def test_eval_forward_matches_reference(name):
    d, p, (x, y_cov, labels), gold, _ = load_case(name)
    with torch.no_grad():
        outs = O.forward(d, p, x, y_cov)
    for k, o in zip(OUT_NAMES, outs):
        assert o.shape == gold["eval_" + k].shape
        assert rel_l2(o, gold["eval_" + k]) < TOL, k


@pytest.mark.parametrize("name", list(CASES))
def test_train_forward_loss_and_grads_match_reference(name):
    d, p, (x, y_cov, labels), gold, full = load_case(name)
    flags = [bool(f) for f in gold["train_flags"]]
    loss, outs, grads = O.loss_and_grads(d, p, x, y_cov, labels, flags)
    assert abs(float(loss) - float(gold["train_loss"])) < 1e-4 * abs(float(gold["train_loss"]))
    for k, o in zip(OUT_NAMES, outs):
        assert rel_l2(o, gold["train_" + k]) < TOL, k
    for pname, g in grads.items():
        if full:
            ref = gold["grad_" + pname]
            assert rel_l2(g, ref) < 2e-4, pname
        else:
            flat = g.reshape(-1).numpy()
            ref = gold["gsample_" + pname]
            assert rel_l2(flat[sample_index(flat.size)], ref) < 2e-4, pname
            assert abs(np.linalg.norm(flat.astype(np.float64)) - gold["gnorm_" + pname]) < 2e-4 * gold["gnorm_" + pname]


def test_coin_flips_follow_numpy_stream():
    d, *_ = load_case("tiny")
    gold = load_case("tiny")[3]
    np.random.seed(7)
    flags = O.draw_teacher_forcing(d, True, int(gold["train_batches_seen"]))
    assert flags == [bool(f) for f in gold["train_flags"]]
    assert O.draw_teacher_forcing(d, False, 0) == [False] * d.horizon


def test_param_shapes_are_reference_state_dict():
    d = O.Dims(num_nodes=207)
    shp = O.param_shapes(d)
    assert len(shp) == 14
    assert shp["encoder.dcrnn_cells.0.gate.weights"] == (390, 128)
    assert shp["decoder.dcrnn_cells.0.update.weights"] == (780, 128)
    assert sum(int(np.prod(s)) for s in shp.values()) == 388761
