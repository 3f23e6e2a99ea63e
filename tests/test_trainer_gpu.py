"""GPU: the trainer's call sequence (model/traintest_MegaCRN.py:101-155) against the drop-in module -- plain torch Adam /
MultiStepLR / loss.backward() / clip_grad_norm_ / evaluate() / torch.save + load_state_dict -- on synthetic arrays of the
METR-LA schema.  The GPU box has no reference checkout, so this is the committed stand-in for running the reference
script itself (megacrn_b200/standin_trainer.py says which reference line each call mirrors); where the checkout exists,
tests/test_launch_traintest.py drives the reference's own file through megacrn_b200.launch_traintest."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MEGACRN_REFERENCE", "/root/reference")


def test_stand_in_trainer_learns_saves_and_reloads(tmp_path):
    from megacrn_b200.standin_trainer import StandInTrainer
    np.random.seed(0)
    torch.manual_seed(0)
    tr = StandInTrainer(num_nodes=207, batch_size=32, n_train=192, n_val=64, n_test=64, workdir=str(tmp_path))
    v0, _ = tr.evaluate(tr.model, "val")
    best, reloaded, (test_loss, test_mae) = tr.fit(epochs=3)
    log = tr.log
    assert len(log) == 3 and tr.batches_seen == 3 * 6
    assert log[-1]["train_loss"] < 0.7 * log[0]["train_loss"], log         # it learns
    assert best < v0, (best, v0)
    assert os.path.exists(tr.path)
    assert abs(reloaded - best) <= 1e-6 * abs(best), (reloaded, best)       # checkpoint round trip: same eval loss
    assert np.isfinite(test_loss) and test_mae > 0
    # the eval fast path (cached prologue) saw every in-place torch.optim update: evaluating twice in a row changes nothing
    a, _ = tr.evaluate(tr.model, "val")
    b, _ = tr.evaluate(tr.model, "val")
    assert a == b


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "model", "traintest_MegaCRN.py")), reason="no reference checkout on this box")
def test_reference_trainer_one_epoch_on_gpu(tmp_path):
    """The reference's own traintest_MegaCRN.py, unmodified, for one epoch against the B200 module (runs only where the
    reference checkout exists)."""
    data = tmp_path / "data"
    data.mkdir()
    from megacrn_b200.standin_trainer import synthetic_npz
    for cat, n, seed in (("train", 128, 0), ("val", 64, 1), ("test", 64, 2)):
        x, y = synthetic_npz(n, 207, seed=seed)
        np.savez(data / f"{cat}.npz", x=x, y=y)
    cmd = [sys.executable, "-m", "megacrn_b200.launch_traintest", "--reference", REF, "--workdir", str(tmp_path / "run"),
           "--data", str(data), "--", "--dataset", "METRLA", "--epochs", "1", "--batch_size", "32"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert "Epoch [1/1]" in out and "Best model performance" in out, out[-3000:]
    assert "Horizon 60mins" in out, out[-3000:]
