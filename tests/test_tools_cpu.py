"""CPU: the data plumbing around the real-data runs of this round (profiles/r2_metrla_real_run.txt, r2_expytky_*.txt):
tools/metrla_npz.py (the reference's window generator restated, HDF5 store read without PyTables / h5py) and the staging
done by megacrn_b200.launch_traintest for the EXPY-TKY harness / the stock-model arm."""
import importlib.util
import os
import struct
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MEGACRN_REFERENCE", "/root/reference")


def _tool():
    spec = importlib.util.spec_from_file_location("_metrla_npz", os.path.join(ROOT, "tools", "metrla_npz.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_windows_follow_the_reference_generator():
    """generate_training_data.py:12-54: x offsets -11..0, y offsets 1..12 around t = 11 .. T-13, features [reading, time of day]."""
    m = _tool()
    T, N = 60, 5
    rng = np.random.default_rng(0)
    values = rng.uniform(0, 70, size=(T, N))
    index = (np.datetime64("2012-03-01T22:00") + np.arange(T) * np.timedelta64(5, "m")).astype("datetime64[ns]")
    x, y, x_off, y_off = m.windows(index, values)
    assert list(x_off) == list(range(-11, 1)) and list(y_off) == list(range(1, 13))
    assert x.shape == (T - 23, 12, N, 2) and y.shape == x.shape
    day = (index - index.astype("datetime64[D]")) / np.timedelta64(1, "D")
    for s in (0, 7, T - 24):
        t = s + 11
        for j, o in enumerate(x_off):
            assert np.array_equal(x[s, j, :, 0], values[t + o]) and np.allclose(x[s, j, :, 1], day[t + o])
        for j, o in enumerate(y_off):
            assert np.array_equal(y[s, j, :, 0], values[t + o]) and np.allclose(y[s, j, :, 1], day[t + o])
    assert 0 <= x[..., 1].min() and x[..., 1].max() < 1          # crosses midnight: still a fraction of a day


def test_contiguous_layout_scan_finds_index_and_values(tmp_path):
    """A minimal byte image with two version-3 contiguous data-layout messages (03 01 <addr> <size>) after the HDF5 signature."""
    m = _tool()
    T, N = 40, 3
    index = (np.datetime64("2012-03-01T00:00") + np.arange(T) * np.timedelta64(5, "m")).astype("datetime64[ns]").astype("<i8")
    values = np.arange(T * N, dtype="<f8").reshape(T, N) % 70
    head = bytearray(b"\x89HDF\r\n\x1a\n" + b"\x00" * 248)
    idx_addr, val_addr = 512, 512 + index.nbytes
    head[64:82] = b"\x03\x01" + struct.pack("<QQ", idx_addr, index.nbytes)
    head[128:146] = b"\x03\x01" + struct.pack("<QQ", val_addr, values.nbytes)
    blob = bytes(head) + b"\x00" * (512 - len(head)) + index.tobytes() + values.tobytes()
    path = tmp_path / "fake.h5"
    path.write_bytes(blob)
    got_index, got_values = m.read_fixed_store(str(path))
    assert got_values.shape == (T, N) and np.array_equal(got_values, values)
    assert np.array_equal(got_index.astype("<i8"), index)
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.h5"
        bad.write_bytes(b"not hdf5" + blob[8:])
        m.read_fixed_store(str(bad))


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "METRLA", "metr-la.h5")), reason="no reference checkout")
def test_real_metrla_store_is_read_completely():
    m = _tool()
    index, values = m.read_fixed_store(os.path.join(REF, "METRLA", "metr-la.h5"))
    assert values.shape == (34272, 207)                       # 119 days of 5-minute readings, 207 sensors
    assert str(index[0])[:16] == "2012-03-01T00:00" and str(index[-1])[:16] == "2012-06-27T23:55"
    assert abs(values.mean() - 53.719) < 1e-3 and values.min() == 0.0 and values.max() <= 70.0


def test_launcher_stages_the_expytky_harness_and_the_stock_model(tmp_path):
    from megacrn_b200 import launch_traintest as L
    ref = tmp_path / "ref"
    (ref / "model_EXPYTKY").mkdir(parents=True)
    (ref / "EXPYTKY").mkdir()
    for name in ("utils.py", "metrics.py", "params.txt", "MegaCRN.py"):
        (ref / "model_EXPYTKY" / name).write_text(f"# reference {name}\n")
    d = L.prepare(str(ref), str(tmp_path / "run"), "EXPYTKY", None, "model_EXPYTKY")
    assert open(os.path.join(d, "MegaCRN.py")).read() == L.SHIM                       # this package's module behind the bare name
    for name in ("utils.py", "metrics.py", "params.txt"):
        assert open(os.path.join(d, name)).read() == f"# reference {name}\n"
    assert os.path.realpath(os.path.join(tmp_path, "run", "EXPYTKY")) == os.path.realpath(ref / "EXPYTKY")
    d2 = L.prepare(str(ref), str(tmp_path / "run_ref"), "EXPYTKY", None, "model_EXPYTKY", stock_model=True)
    assert open(os.path.join(d2, "MegaCRN.py")).read() == "# reference MegaCRN.py\n"  # baseline arm: the reference's own model file


def test_dataframe_values_are_writable_for_the_reference_utils():
    """model_EXPYTKY/utils.py:56-57 clips `df[...].values` in place; pandas >= 3 returns read-only arrays."""
    pd = pytest.importorskip("pandas")
    from megacrn_b200 import launch_traintest as L
    L._writable_dataframe_values()
    v = pd.DataFrame({"speed": [-1.0, 50.0, 250.0]})[["speed"]].values
    v[v < 0] = 0
    v[v > 200.0] = 100.0
    assert v.ravel().tolist() == [0.0, 50.0, 100.0]
