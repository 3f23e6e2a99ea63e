"""Run the reference's own ``traintest_MegaCRN.py`` UNCHANGED against the B200 module.

    python -m megacrn_b200.launch_traintest --reference /path/to/MegaCRN [--workdir run] [--data DIR] \\
           -- --dataset METRLA --gpu 0 --epochs 200

The reference script (model/traintest_MegaCRN.py) imports ``MegaCRN`` and ``utils`` by bare name (:14-15), writes
``../save/...`` relative to the CWD (:201-206), copies ``MegaCRN.py`` / ``utils.py`` from the CWD into the run directory
(:207-209) and reads ``../<DATASET>/{train,val,test}.npz`` (:271).  This launcher therefore

  * creates ``<workdir>/model/`` and makes it the CWD, holding a one-line shim ``MegaCRN.py`` that re-exports
    ``megacrn_b200.MegaCRN`` and a copy of the reference's ``utils.py`` (data plumbing: out of the hot path, used as is);
  * links ``<workdir>/<DATASET>`` to the dataset directory (``--data`` or ``<reference>/<DATASET>``);
  * provides a ``torchsummary`` stub if that package is absent (imported at :11, never used by the script);
  * executes the reference file with ``runpy.run_path(..., run_name="__main__")`` with the shim directory first on
    ``sys.path``.  Nothing of the reference is modified or copied into this repository.

``--stock-model`` swaps the shim for the reference's own ``MegaCRN.py`` (the unmodified reference end to end: the baseline
arm of a side-by-side run) and ``--seed`` seeds torch / numpy / random first, so that both arms start from the same weights
and draw the same batches and teacher-forcing coins.

``--harness model_EXPYTKY`` runs the EXPY-TKY harness (model_EXPYTKY/traintest_MegaCRN.py) the same way: it also reads
``params.txt`` from the CWD (:182) and copies ``metrics.py`` / ``params.txt`` next to the model (:205-206), its data directory
is always ``../EXPYTKY`` (params.txt), it calls ``torchsummary.summary`` (:27, stubbed to a no-op when the package is absent)
and its ``utils.py`` imports ``jpholiday`` for a helper the trainer never calls (utils.py:5, :113; stubbed when absent).
"""
from __future__ import annotations

import argparse
import os
import runpy
import shutil
import sys
import types

SHIM = '"""Shim written by megacrn_b200.launch_traintest: the reference trainer imports MegaCRN by bare name."""\n' \
       "from megacrn_b200.MegaCRN import *  # noqa: F401,F403\n" \
       "from megacrn_b200.MegaCRN import MegaCRN, print_params  # noqa: F401\n"


# plumbing files each harness expects in its CWD (data handling / metrics / the month table: out of the hot path, used as is)
HARNESS_FILES = {"model": ("utils.py",), "model_EXPYTKY": ("utils.py", "metrics.py", "params.txt")}


def prepare(reference: str, workdir: str, dataset: str, data: str | None, harness: str = "model",
            stock_model: bool = False) -> str:
    model_dir = os.path.join(workdir, harness)
    os.makedirs(model_dir, exist_ok=True)
    if stock_model:          # baseline arm: the reference's own model file, i.e. the unmodified reference end to end
        shutil.copy2(os.path.join(reference, harness, "MegaCRN.py"), os.path.join(model_dir, "MegaCRN.py"))
    else:
        with open(os.path.join(model_dir, "MegaCRN.py"), "w") as f:
            f.write(SHIM)
    for name in HARNESS_FILES[harness]:
        shutil.copy2(os.path.join(reference, harness, name), os.path.join(model_dir, name))
    src = os.path.abspath(data or os.path.join(reference, dataset))
    dst = os.path.join(workdir, dataset)
    if not os.path.exists(dst):
        os.symlink(src, dst, target_is_directory=True)
    return model_dir


def _writable_dataframe_values():
    """pandas >= 3 hands out read-only arrays from ``DataFrame.values`` (copy-on-write); the EXPY-TKY ``utils.get_data``
    clips them in place (model_EXPYTKY/utils.py:56-57, written for pandas 1.x).  Return a writable copy instead."""
    try:
        import pandas as pd
    except ImportError:
        return
    if int(pd.__version__.split(".")[0]) < 3:
        return
    orig = pd.DataFrame.values

    def values(self):
        v = orig.fget(self)
        return v if v.flags.writeable else v.copy()
    pd.DataFrame.values = property(values, doc=orig.__doc__)


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", required=True, help="checkout of deepkashiwa20/MegaCRN")
    ap.add_argument("--workdir", default="run")
    ap.add_argument("--data", default=None, help="directory holding train/val/test.npz (default <reference>/<DATASET>)")
    ap.add_argument("--script", default="traintest_MegaCRN.py")
    ap.add_argument("--harness", default="model", choices=sorted(HARNESS_FILES), help="reference directory holding the trainer")
    ap.add_argument("--stock-model", action="store_true",
                    help="baseline arm: run the harness against the reference's OWN MegaCRN.py (nothing of this package on the path)")
    ap.add_argument("--seed", type=int, default=None,
                    help="seed torch / numpy / random before the script starts (the reference scripts leave seeding commented "
                         "out): two arms with the same seed see the same initial weights, batches and teacher-forcing coins")
    ap.add_argument("rest", nargs=argparse.REMAINDER, help="arguments after -- go to the reference script")
    args = ap.parse_args(argv)
    rest = [a for a in args.rest if a != "--"]
    dataset = "METRLA"
    if "--dataset" in rest:
        dataset = rest[rest.index("--dataset") + 1]
    if args.harness == "model_EXPYTKY":
        dataset = "EXPYTKY"                    # 'EXPYTKY' and 'EXPYTKY*' both live in ../EXPYTKY (params.txt)
    workdir = os.path.abspath(args.workdir)
    model_dir = prepare(os.path.abspath(args.reference), workdir, dataset, args.data, args.harness, args.stock_model)
    script = os.path.join(os.path.abspath(args.reference), args.harness, args.script)
    try:
        import torchsummary  # noqa: F401
    except ImportError:
        stub = types.ModuleType("torchsummary")
        stub.summary = lambda *a, **k: None
        sys.modules["torchsummary"] = stub
    try:
        import jpholiday  # noqa: F401
    except ImportError:
        stub = types.ModuleType("jpholiday")
        stub.is_holiday = lambda *a, **k: False
        sys.modules["jpholiday"] = stub
    _writable_dataframe_values()
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (repo_root, model_dir):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    os.chdir(model_dir)
    if args.seed is not None:
        import random

        import numpy as np
        import torch
        random.seed(args.seed)
        np.random.seed(args.seed)
        torch.manual_seed(args.seed)
    sys.argv = [script] + rest
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
