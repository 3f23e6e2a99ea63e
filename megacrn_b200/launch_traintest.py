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


def prepare(reference: str, workdir: str, dataset: str, data: str | None) -> str:
    model_dir = os.path.join(workdir, "model")
    os.makedirs(model_dir, exist_ok=True)
    with open(os.path.join(model_dir, "MegaCRN.py"), "w") as f:
        f.write(SHIM)
    shutil.copy2(os.path.join(reference, "model", "utils.py"), os.path.join(model_dir, "utils.py"))
    src = os.path.abspath(data or os.path.join(reference, dataset))
    dst = os.path.join(workdir, dataset)
    if not os.path.exists(dst):
        os.symlink(src, dst, target_is_directory=True)
    return model_dir


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", required=True, help="checkout of deepkashiwa20/MegaCRN")
    ap.add_argument("--workdir", default="run")
    ap.add_argument("--data", default=None, help="directory holding train/val/test.npz (default <reference>/<DATASET>)")
    ap.add_argument("--script", default="traintest_MegaCRN.py")
    ap.add_argument("rest", nargs=argparse.REMAINDER, help="arguments after -- go to the reference script")
    args = ap.parse_args(argv)
    rest = [a for a in args.rest if a != "--"]
    dataset = "METRLA"
    if "--dataset" in rest:
        dataset = rest[rest.index("--dataset") + 1]
    workdir = os.path.abspath(args.workdir)
    model_dir = prepare(os.path.abspath(args.reference), workdir, dataset, args.data)
    script = os.path.join(os.path.abspath(args.reference), "model", args.script)
    try:
        import torchsummary  # noqa: F401
    except ImportError:
        stub = types.ModuleType("torchsummary")
        stub.summary = lambda *a, **k: None
        sys.modules["torchsummary"] = stub
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (repo_root, model_dir):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    os.chdir(model_dir)
    sys.argv = [script] + rest
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
