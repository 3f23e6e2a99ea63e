"""CPU: the reference trainer, unmodified, resolves `from MegaCRN import MegaCRN` to this repo's module when started
through megacrn_b200.launch_traintest.  Without a GPU the run must stop at the module's explicit no-CPU-path error --
which proves the wiring (and that there is no silent fallback).  Needs the reference checkout; on a GPU box the same
launcher is run for a full epoch by tests/test_trainer_gpu.py::test_reference_trainer_one_epoch_on_gpu."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

REF = os.environ.get("MEGACRN_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "model", "traintest_MegaCRN.py")), reason="no reference checkout")
@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_reference_trainer_runs_against_the_b200_module(tmp_path):
    data = tmp_path / "data"
    data.mkdir()
    rng = np.random.default_rng(0)
    for cat, n in (("train", 8), ("val", 4), ("test", 4)):
        x = rng.normal(size=(n, 12, 207, 2))
        y = rng.normal(size=(n, 12, 207, 2))
        np.savez(data / f"{cat}.npz", x=x, y=y)
    cmd = [sys.executable, "-m", "megacrn_b200.launch_traintest", "--reference", REF, "--workdir", str(tmp_path / "run"),
           "--data", str(data), "--", "--dataset", "METRLA", "--epochs", "1", "--batch_size", "4"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    out = r.stdout + r.stderr
    assert "Trainable parameter list" in out or "trainable parameters" in out, out[-2000:]     # reference print_model ran
    assert "388761" in out, out[-2000:]                                                        # our module, reference shapes
    assert "megacrn_b200.MegaCRN has no CPU path" in out, out[-2000:]
    assert os.path.exists(tmp_path / "run" / "model" / "MegaCRN.py")
