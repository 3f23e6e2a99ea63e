"""CPU: the drop-in module registers the reference's 14 state_dict entries (names, order, shapes) and, built under the same
torch seed, holds bit-identical initial weights (model/MegaCRN.py:8-14, :130-144, :149-157) -- so a seeded reference run and
a seeded run on this module start from the same point.  Needs the reference checkout (build container); the key / shape
half runs everywhere."""
import importlib.util
import os

import pytest
import torch

REF = os.environ.get("MEGACRN_REFERENCE", "/root/reference")
REF_MODEL = os.path.join(REF, "model", "MegaCRN.py")

CASES = [dict(num_nodes=207, input_dim=1, output_dim=1, horizon=12, rnn_units=64),
         dict(num_nodes=33, input_dim=2, output_dim=1, horizon=4, rnn_units=16, mem_num=6, mem_dim=12, cheb_k=2),
         dict(num_nodes=21, input_dim=1, output_dim=1, horizon=3, rnn_units=12, mem_num=5, mem_dim=8, num_layers=3)]


def _reference_class():
    spec = importlib.util.spec_from_file_location("_reference_megacrn", REF_MODEL)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MegaCRN


def test_state_dict_keys_and_shapes_follow_the_oracle_table():
    from megacrn_b200 import MegaCRN, _abi
    from oracle import megacrn_oracle as O
    for kw in CASES:
        m = MegaCRN(**kw)
        d = O.Dims(**kw)
        shapes = O.param_shapes(d)
        sd = m.state_dict()
        layers = kw.get("num_layers", 1)
        assert list(sd.keys()) == list(shapes.keys())
        if layers == 1:
            assert list(sd.keys()) == list(_abi.STATE_DICT_KEYS)
        # C-ABI order: the 14 tensors of mcrn_params, then 8 per stacked layer (mcrn_layer_params)
        abi_keys = _abi.param_keys(layers)
        assert abi_keys[:14] == _abi.STATE_DICT_KEYS and sorted(abi_keys) == sorted(sd.keys())
        named = dict(m.named_parameters())
        assert [id(named[k]) for k in abi_keys] == [id(t) for t in m._ordered_params()]
        for k, v in sd.items():
            assert tuple(v.shape) == tuple(shapes[k]), k


@pytest.mark.skipif(not os.path.exists(REF_MODEL), reason="no reference checkout on this machine")
@pytest.mark.parametrize("kw", CASES)
def test_same_seed_gives_the_reference_initial_weights(kw):
    from megacrn_b200 import MegaCRN
    Ref = _reference_class()
    torch.manual_seed(1234)
    ref = Ref(**kw)
    torch.manual_seed(1234)
    ours = MegaCRN(**kw)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k
    # the memory bank of construct_memory (:149-157) in particular
    for k in ("memory.Memory", "memory.Wq", "memory.We1", "memory.We2"):
        assert a[k].abs().sum() > 0 and torch.equal(a[k], b[k])
