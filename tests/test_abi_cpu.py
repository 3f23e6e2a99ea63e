"""CPU: the C-ABI library loads, exports every symbol include/megacrn_b200.h declares, validates its
arguments, and refuses to compute without an sm_100 device (no fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="megacrn_b200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcrn_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from megacrn_b200 import _abi
    lib = _abi.load()
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.mcrn_abi_version() == 1
    # the public header holds the drop-in boundary only; debug / probe / tuning entries live in megacrn_b200_debug.h
    assert not [s for s in syms if "debug" in s or "probe" in s or s in ("mcrn_set_option", "mcrn_kernel_timing")], syms
    dbg = _declared_symbols("megacrn_b200_debug.h")
    assert "mcrn_debug_fused_timeline" in dbg and "mcrn_set_option" in dbg
    for s in dbg:
        assert hasattr(lib, s), s


def test_workspace_and_dim_validation():
    from megacrn_b200 import _abi
    lib = _abi.load()
    d = _abi.Dims(batch=64, num_nodes=207, seq_len=12, horizon=12, input_dim=1, output_dim=1, ycov_dim=1,
                  rnn_units=64, num_layers=1, cheb_k=3, mem_num=20, mem_dim=64)
    fwd = lib.mcrn_workspace_bytes(d, 0)
    trn = lib.mcrn_workspace_bytes(d, 1)
    assert 0 < fwd < trn
    assert trn > 12 * 5 * 207 * 64 * (64 + 128) * 4          # per-step XP buffers are saved
    assert lib.mcrn_host_workspace_bytes(d, 0) > fwd
    d.num_layers = 2                                         # stacked cells: more buffers, still valid
    assert lib.mcrn_workspace_bytes(d, 1) > trn
    d.num_layers = 5
    assert lib.mcrn_workspace_bytes(d, 0) == 0
    assert b"num_layers" in lib.mcrn_last_error()
    d.num_layers, d.cheb_k = 1, 1
    assert lib.mcrn_workspace_bytes(d, 0) == 0
    assert lib.mcrn_support_ld(207) == 208 and lib.mcrn_support_ld(208) == 208


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_device_means_error_not_fallback():
    from megacrn_b200 import _abi
    lib = _abi.load()
    assert lib.mcrn_device_ok() == -4
    a = (C.c_float * 16)()
    st = lib.mcrn_gemm(4, 4, 4, C.addressof(a), 4, 0, C.addressof(a), 4, 0, C.addressof(a), 4, 0, None)
    assert st == -4 and b"CUDA" in lib.mcrn_last_error() or b"device" in lib.mcrn_last_error()


def test_module_keeps_reference_interface():
    from megacrn_b200 import MegaCRN, _abi
    m = MegaCRN(num_nodes=20, input_dim=1, output_dim=1, horizon=3, rnn_units=8, num_layers=1, mem_num=4, mem_dim=8,
                cheb_k=3, cl_decay_steps=2000, use_curriculum_learning=True)
    assert tuple(m.state_dict().keys()) == _abi.STATE_DICT_KEYS
    assert abs(m.compute_sampling_threshold(0) - 2000 / 2001) < 1e-12
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 2, 20, 1), torch.zeros(1, 3, 20, 1))


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """The boundary is a C ABI: include/megacrn_b200.h compiles as C99 (no C++ / torch types) and a C program links against
    libmegacrn_b200.so and calls it (argument validation only -- no GPU here)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "host.c"
    src.write_text(r'''
#include <stdio.h>
#include "megacrn_b200.h"
int main(void) {
  mcrn_dims d = {64, 207, 12, 12, 1, 1, 1, 64, 1, 3, 20, 64};
  size_t fwd = mcrn_workspace_bytes(&d, 0), trn = mcrn_workspace_bytes(&d, MCRN_FWD_SAVE_FOR_BACKWARD);
  d.num_layers = 9;
  size_t bad = mcrn_workspace_bytes(&d, 0);
  printf("%d %zu %zu %zu %d\n", mcrn_abi_version(), fwd, trn, bad, mcrn_support_ld(207));
  return (mcrn_abi_version() == 1 && fwd > 0 && trn > fwd && bad == 0) ? 0 : 1;
}
''')
    libdir = os.path.join(ROOT, "megacrn_b200")
    exe = tmp_path / "host"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:libmegacrn_b200.so", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.split()[-1] == "208"
