"""ctypes binding of libmegacrn_b200.so (C ABI declared in include/megacrn_b200.h).

There is no fallback: if the shared library is missing or the device is not an
sm_100 part, every compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmegacrn_b200.so")

MCRN_FWD_SAVE_FOR_BACKWARD = 1
MCRN_FWD_REUSE_PROLOGUE = 2
ABI_VERSION = 1

PARAM_FIELDS = (
    "memory", "wq", "we1", "we2",
    "enc_gate_w", "enc_gate_b", "enc_update_w", "enc_update_b",
    "dec_gate_w", "dec_gate_b", "dec_update_w", "dec_update_b",
    "proj_w", "proj_b",
)
# state_dict key of each field, in the reference's registration order
STATE_DICT_KEYS = (
    "memory.Memory", "memory.Wq", "memory.We1", "memory.We2",
    "encoder.dcrnn_cells.0.gate.weights", "encoder.dcrnn_cells.0.gate.bias",
    "encoder.dcrnn_cells.0.update.weights", "encoder.dcrnn_cells.0.update.bias",
    "decoder.dcrnn_cells.0.gate.weights", "decoder.dcrnn_cells.0.gate.bias",
    "decoder.dcrnn_cells.0.update.weights", "decoder.dcrnn_cells.0.update.bias",
    "proj.0.weight", "proj.0.bias",
)


# stacked cells of one layer i >= 1 (mcrn_layer_params): field -> state_dict key template
LAYER_FIELDS = (
    "enc_gate_w", "enc_gate_b", "enc_update_w", "enc_update_b",
    "dec_gate_w", "dec_gate_b", "dec_update_w", "dec_update_b",
)
LAYER_KEYS = (
    "encoder.dcrnn_cells.{i}.gate.weights", "encoder.dcrnn_cells.{i}.gate.bias",
    "encoder.dcrnn_cells.{i}.update.weights", "encoder.dcrnn_cells.{i}.update.bias",
    "decoder.dcrnn_cells.{i}.gate.weights", "decoder.dcrnn_cells.{i}.gate.bias",
    "decoder.dcrnn_cells.{i}.update.weights", "decoder.dcrnn_cells.{i}.update.bias",
)
MAX_LAYERS = 4


def param_keys(num_layers: int = 1):
    """state_dict keys in the order the C ABI takes the tensors: the 14 of mcrn_params, then 8 per stacked layer."""
    keys = list(STATE_DICT_KEYS)
    for i in range(1, num_layers):
        keys += [k.format(i=i) for k in LAYER_KEYS]
    return tuple(keys)


class Dims(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "batch", "num_nodes", "seq_len", "horizon", "input_dim", "output_dim", "ycov_dim",
        "rnn_units", "num_layers", "cheb_k", "mem_num", "mem_dim")]


class Params(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in PARAM_FIELDS]


class LayerParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in LAYER_FIELDS]


class MegaCRNLibraryError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MegaCRNLibraryError(
            f"{LIB_PATH} is missing: build it with `make -C megacrn_b200/csrc` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). "
            "megacrn_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, u8p, fp = C.c_void_p, C.c_char_p, C.c_void_p
    lib.mcrn_abi_version.restype = C.c_int
    lib.mcrn_last_error.restype = C.c_char_p
    lib.mcrn_device_ok.restype = C.c_int
    lib.mcrn_launch_count.restype = C.c_uint64
    lib.mcrn_set_engine.argtypes = [C.c_int]
    lib.mcrn_get_engine.restype = C.c_int
    lib.mcrn_mode_epoch.restype = C.c_uint64
    lib.mcrn_debug_probe_mn16.restype = C.c_int
    lib.mcrn_debug_probe_mn16.argtypes = [vp, vp, vp, C.c_uint, C.c_uint, C.c_uint, C.c_uint, C.c_uint, vp]
    lib.mcrn_set_option.restype = C.c_int
    lib.mcrn_set_option.argtypes = [C.c_char_p, C.c_int]
    lib.mcrn_debug_fused_timeline.restype = C.c_int
    lib.mcrn_debug_fused_timeline.argtypes = [C.c_void_p, C.c_int]
    lib.mcrn_debug_launch_spans.restype = C.c_int
    lib.mcrn_debug_launch_spans.argtypes = [C.c_void_p, C.c_int]
    lib.mcrn_adam_step.restype = C.c_int
    lib.mcrn_adam_step.argtypes = [C.POINTER(Dims), C.POINTER(Params), C.POINTER(Params), C.POINTER(Params), C.POINTER(Params),
                                   fp, C.c_float, C.c_float, C.c_float, C.c_float, vp]
    lib.mcrn_adam_step_layers.restype = C.c_int
    lib.mcrn_adam_step_layers.argtypes = [C.POINTER(Dims), C.POINTER(Params), vp, C.POINTER(Params), vp, C.POINTER(Params), vp,
                                          C.POINTER(Params), vp, fp, C.c_float, C.c_float, C.c_float, C.c_float, vp]
    lib.mcrn_kernel_timing.restype = C.c_int
    lib.mcrn_kernel_timing.argtypes = [C.c_int]
    lib.mcrn_kernel_timing_read.restype = C.c_int
    lib.mcrn_kernel_timing_read.argtypes = [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    lib.mcrn_set_bwd_fused.restype = C.c_int
    lib.mcrn_set_bwd_fused.argtypes = [C.c_int]
    lib.mcrn_set_fused.restype = C.c_int
    lib.mcrn_set_fused.argtypes = [C.c_int, C.c_int]
    lib.mcrn_support_ld.argtypes = [C.c_int]
    lib.mcrn_workspace_bytes.restype = C.c_size_t
    lib.mcrn_workspace_bytes.argtypes = [C.POINTER(Dims), C.c_uint32]
    lib.mcrn_forward.restype = C.c_int
    lib.mcrn_forward.argtypes = [C.POINTER(Dims), C.POINTER(Params), fp, fp, fp, u8p,
                                 fp, fp, fp, fp, fp, vp, C.c_size_t, C.c_uint32, vp]
    lib.mcrn_backward.restype = C.c_int
    lib.mcrn_backward.argtypes = [C.POINTER(Dims), C.POINTER(Params), fp, fp, fp, u8p,
                                  fp, fp, fp, fp, fp, C.POINTER(Params), vp, C.c_size_t, vp]
    lib.mcrn_forward_layers.restype = C.c_int
    lib.mcrn_forward_layers.argtypes = [C.POINTER(Dims), C.POINTER(Params), vp, fp, fp, fp, u8p,
                                        fp, fp, fp, fp, fp, vp, C.c_size_t, C.c_uint32, vp]
    lib.mcrn_backward_layers.restype = C.c_int
    lib.mcrn_backward_layers.argtypes = [C.POINTER(Dims), C.POINTER(Params), vp, fp, fp, fp, u8p,
                                         fp, fp, fp, fp, fp, C.POINTER(Params), vp, vp, C.c_size_t, vp]
    lib.mcrn_trainer_loss.restype = C.c_int
    lib.mcrn_trainer_loss.argtypes = [C.POINTER(Dims), fp, fp, fp, fp, fp, C.c_float, C.c_float, C.c_float,
                                      C.c_float, fp, fp, fp, vp, C.c_size_t, vp]
    lib.mcrn_trainer_loss_dp.restype = C.c_int
    lib.mcrn_trainer_loss_dp.argtypes = [C.POINTER(Dims), fp, fp, fp, fp, fp, C.c_float, C.c_float, C.c_float,
                                         C.c_float, fp, fp, fp, fp, vp, C.c_size_t, vp]
    lib.mcrn_mask_count.restype = C.c_int
    lib.mcrn_mask_count.argtypes = [fp, C.c_int64, C.c_float, C.c_float, fp, vp]
    lib.mcrn_supports_fwd.restype = C.c_int
    lib.mcrn_supports_fwd.argtypes = [C.POINTER(Dims), fp, fp, fp, fp, vp, C.c_size_t, vp]
    lib.mcrn_supports_fwd2.restype = C.c_int
    lib.mcrn_supports_fwd2.argtypes = [C.POINTER(Dims), fp, fp, fp, fp, fp, vp, C.c_size_t, vp]
    lib.mcrn_gemm.restype = C.c_int
    lib.mcrn_gemm.argtypes = [C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int, fp, C.c_int, C.c_int,
                              fp, C.c_int, C.c_int, vp]
    lib.mcrn_debug_tc_gemm.restype = C.c_int
    lib.mcrn_debug_tc_gemm.argtypes = [C.c_int, C.c_int, C.c_int, fp, C.c_int, C.c_int, fp, C.c_int, C.c_int,
                                       fp, C.c_int, fp, vp]
    lib.mcrn_host_workspace_bytes.restype = C.c_size_t
    lib.mcrn_host_workspace_bytes.argtypes = [C.POINTER(Dims), C.c_uint32]
    lib.mcrn_forward_host.restype = C.c_int
    lib.mcrn_forward_host.argtypes = [C.POINTER(Dims), C.POINTER(Params), fp, fp, fp, u8p,
                                      fp, fp, fp, fp, fp, vp, C.c_size_t, C.c_uint32, vp]
    if lib.mcrn_abi_version() != ABI_VERSION:
        raise MegaCRNLibraryError(f"ABI mismatch: library {lib.mcrn_abi_version()} vs binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().mcrn_last_error().decode("utf-8", "replace")
        raise MegaCRNLibraryError(f"{what} failed with status {status}: {msg}")


def ptr(t) -> Optional[int]:
    return None if t is None else t.data_ptr()


def make_params(tensors: Sequence) -> Params:
    p = Params()
    for name, t in zip(PARAM_FIELDS, tensors):
        setattr(p, name, t.data_ptr())
    return p


def make_layer_params(tensors: Sequence):
    """``tensors``: 8 per stacked layer in LAYER_FIELDS order -> a ctypes array of mcrn_layer_params (None if empty).
    Keep the returned object alive across the call that takes ``C.addressof`` / ``C.byref`` of it."""
    n = len(tensors) // len(LAYER_FIELDS)
    assert n * len(LAYER_FIELDS) == len(tensors) and n <= MAX_LAYERS - 1
    if n == 0:
        return None
    arr = (LayerParams * n)()
    for l in range(n):
        for j, name in enumerate(LAYER_FIELDS):
            setattr(arr[l], name, tensors[l * len(LAYER_FIELDS) + j].data_ptr())
    return arr


def layer_ptr(arr) -> Optional[int]:
    return None if arr is None else C.addressof(arr)


def tf_bytes(flags: Optional[Sequence[bool]], horizon: int) -> bytes:
    if flags is None:
        return bytes(horizon)
    assert len(flags) == horizon
    return bytes(1 if f else 0 for f in flags)
