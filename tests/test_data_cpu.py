"""CPU: the device-resident input pipeline (megacrn_b200/data.py) reproduces the reference DataLoader + prepare_x_y
(model/utils.py:6-43, model/traintest_MegaCRN.py:33-48; restated in oracle/megacrn_oracle.py:DataLoaderOracle and
pinned here against the reference module itself when /root/reference is present)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from megacrn_b200.data import DeviceDataLoader
from oracle.megacrn_oracle import DataLoaderOracle


def _data(n=23, t=4, nodes=5):
    g = np.random.default_rng(0)
    return g.standard_normal((n, t, nodes, 2)), g.standard_normal((n, t, nodes, 2))


@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("bs,pad", [(8, True), (8, False), (23, True), (5, True)])
def test_device_loader_matches_oracle(bs, pad, shuffle):
    xs, ys = _data()
    np.random.seed(11)
    ref = DataLoaderOracle(xs, ys, bs, pad_with_last_sample=pad, shuffle=shuffle)
    after_ref = np.random.uniform()
    np.random.seed(11)
    dl = DeviceDataLoader(xs, ys, bs, pad_with_last_sample=pad, shuffle=shuffle, device="cpu")
    assert np.random.uniform() == after_ref                    # same NumPy stream consumption
    assert (dl.size, dl.num_batch) == (ref.size, ref.num_batch)
    got = list(dl.get_iterator())
    want = list(ref.batches())
    assert len(got) == len(want) == ref.num_batch
    for (x, y, c), (rx, ry, rc) in zip(got, want):
        assert x.dtype == torch.float32 and x.shape == rx.shape
        np.testing.assert_array_equal(x.numpy(), rx)
        np.testing.assert_array_equal(y.numpy(), ry)
        np.testing.assert_array_equal(c.numpy(), rc)


@pytest.mark.skipif(not os.path.exists("/root/reference/model/utils.py"), reason="reference checkout not present")
def test_oracle_loader_matches_reference_module():
    spec = importlib.util.spec_from_file_location("ref_utils", "/root/reference/model/utils.py")
    ref_utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_utils)
    xs, ys = _data()
    for bs, pad, shuffle in [(8, True, False), (8, True, True), (5, False, True)]:
        np.random.seed(3)
        r = ref_utils.DataLoader(xs, ys, bs, pad_with_last_sample=pad, shuffle=shuffle)
        np.random.seed(3)
        o = DataLoaderOracle(xs, ys, bs, pad_with_last_sample=pad, shuffle=shuffle)
        assert (r.size, r.num_batch) == (o.size, o.num_batch)
        for (rx, ry), (ox, oy, oc) in zip(r.get_iterator(), o.batches()):
            np.testing.assert_array_equal(rx[..., :1].astype(np.float32), ox)
            np.testing.assert_array_equal(ry[..., :1].astype(np.float32), oy)
            np.testing.assert_array_equal(ry[..., 1:].astype(np.float32), oc)
