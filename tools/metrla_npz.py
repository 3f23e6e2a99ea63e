"""Build the METR-LA / PEMS-BAY `train.npz`, `val.npz`, `test.npz` the reference trainer reads (model/traintest_MegaCRN.py:269-273)
from the raw `metr-la.h5`, without PyTables / h5py (neither is in this image).

The file is a pandas "fixed"-format HDF5 store: the readings are ONE contiguous, uncompressed float64 dataset
`/df/block0_values` [T][N] and the timestamps one contiguous int64 dataset `/df/axis1` [T] (ns since the epoch).  Both are
located through their HDF5 data-layout messages (version 3, class 1 = contiguous: `03 01 <address u64> <size u64>`): the
largest such extent is the value block, the one 1/N of its size with regular 5-minute increments is the index.  Everything
is checked (regular timestamps, value range) before anything is written.

The windows follow the reference's generate_training_data.py (:57-103): 12 input offsets -11..0, 12 target offsets 1..12,
features [reading, time of day as a fraction of a day], the last 20 % of the windows for testing, the first 70 % for
training, the rest for validation; float64 arrays `x`, `y` (+ the offset vectors) per file.

    python tools/metrla_npz.py --h5 <dir>/metr-la.h5 --out <dir>
"""
import argparse
import os
import struct

import numpy as np


def contiguous_extents(buf: bytes, header_bytes: int = 1 << 20):
    """(address, size) of every contiguous-layout message found in the first `header_bytes` of the file."""
    out, s = [], 0
    head = buf[:header_bytes]
    while True:
        i = head.find(b"\x03\x01", s)
        if i < 0:
            break
        s = i + 1
        if i + 18 > len(head):
            continue
        addr, size = struct.unpack("<QQ", head[i + 2:i + 18])
        if 0 < addr < len(buf) and 0 < size <= len(buf) - addr:
            out.append((addr, size))
    return out


def read_fixed_store(path):
    buf = open(path, "rb").read()
    if buf[:8] != b"\x89HDF\r\n\x1a\n":
        raise ValueError(f"{path}: not an HDF5 file")
    ext = contiguous_extents(buf)
    if not ext:
        raise ValueError("no contiguous dataset found (compressed / chunked store?)")
    vaddr, vsize = max(ext, key=lambda e: e[1])
    index = None
    for addr, size in ext:
        if size % 8 or vsize % size or size == vsize:
            continue
        t = np.frombuffer(buf, dtype="<i8", count=size // 8, offset=addr)
        d = np.diff(t)
        if len(d) and d.min() == d.max() and d[0] > 0:
            index = t
            break
    if index is None:
        raise ValueError("no regular int64 time index matching the value block")
    T = index.size
    N = vsize // 8 // T
    values = np.frombuffer(buf, dtype="<f8", count=T * N, offset=vaddr).reshape(T, N)
    if not np.isfinite(values).all() or values.min() < 0 or values.max() > 200:
        raise ValueError("value block failed the range check")
    return index.astype("datetime64[ns]"), values


def windows(index, values):
    T, N = values.shape
    day = (index - index.astype("datetime64[D]")) / np.timedelta64(1, "D")            # time of day in [0, 1)
    data = np.stack([values, np.broadcast_to(day[:, None], (T, N))], axis=-1)         # [T][N][2]
    x_off, y_off = np.arange(-11, 1), np.arange(1, 13)
    first, stop = 11, T - 12                                                         # t = index of the last observation
    win = np.lib.stride_tricks.sliding_window_view(data, 24, axis=0)                  # [T-23][N][2][24]
    win = win[:stop - first].transpose(0, 3, 1, 2)                                    # [S][24][N][2]
    return np.ascontiguousarray(win[:, :12]), np.ascontiguousarray(win[:, 12:]), x_off, y_off


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--h5", required=True)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    index, values = read_fixed_store(args.h5)
    print("readings", values.shape, "from", index[0], "to", index[-1], "mean %.3f" % values.mean())
    x, y, x_off, y_off = windows(index, values)
    S = x.shape[0]
    n_test, n_train = round(S * 0.2), round(S * 0.7)
    n_val = S - n_test - n_train
    parts = {"train": slice(0, n_train), "val": slice(n_train, n_train + n_val), "test": slice(S - n_test, S)}
    os.makedirs(args.out, exist_ok=True)
    for name, sl in parts.items():
        print(name, "x:", x[sl].shape, "y:", y[sl].shape)
        np.savez(os.path.join(args.out, name + ".npz"), x=x[sl], y=y[sl], x_offsets=x_off[:, None], y_offsets=y_off[:, None])


if __name__ == "__main__":
    main()
