"""Device-resident input pipeline.

The reference keeps the dataset in host numpy arrays and, every step, slices a batch, converts it with
``torch.from_numpy(...).float()`` and copies three tensors to the device (model/utils.py:6-43 ``DataLoader``,
model/traintest_MegaCRN.py:33-48 ``prepare_x_y``).  Here the (already scaled) dataset is uploaded ONCE as fp32; batches are
device-side slices.  Same semantics as the reference loader: pad with the last sample to a multiple of ``batch_size``,
one ``np.random.permutation`` at construction when ``shuffle`` (same NumPy stream consumption), ``num_batch`` full batches.
"""
from __future__ import annotations

import numpy as np
import torch


class DeviceDataLoader:
    """Drop-in for the reference ``DataLoader`` whose iterator yields what ``prepare_x_y`` would return.

    ``get_iterator()`` yields ``(x, y, y_cov)`` device tensors: ``x = xs[..., :input_dim]``, ``y = ys[..., :output_dim]``,
    ``y_cov = ys[..., output_dim:]`` (model/traintest_MegaCRN.py:41-47), fp32, already on ``device``.
    ``get_raw_iterator()`` yields the reference's ``(x_i, y_i)`` pairs (full channel width) as device tensors."""

    def __init__(self, xs, ys, batch_size, pad_with_last_sample=True, shuffle=False, device="cuda", input_dim=1,
                 output_dim=1):
        xs, ys = np.asarray(xs), np.asarray(ys)
        self.batch_size = batch_size
        self.current_ind = 0
        n = len(xs)
        index = np.arange(n)
        if pad_with_last_sample:                                   # utils.py:17-22
            num_padding = (batch_size - (n % batch_size)) % batch_size
            index = np.concatenate([index, np.full(num_padding, n - 1, dtype=index.dtype)])
        self.size = len(index)
        self.num_batch = int(self.size // self.batch_size)         # utils.py:23-24
        if shuffle:                                                # utils.py:25-27: one permutation, drawn from np.random
            index = index[np.random.permutation(self.size)]
        self.input_dim, self.output_dim = input_dim, output_dim
        dev = torch.device(device)
        idx = torch.from_numpy(index).to(dev)
        # one upload of the whole (scaled) dataset; the permutation / padding is applied on the device
        self.xs = torch.from_numpy(np.ascontiguousarray(xs)).to(dev, dtype=torch.float32)[idx].contiguous()
        self.ys = torch.from_numpy(np.ascontiguousarray(ys)).to(dev, dtype=torch.float32)[idx].contiguous()

    def _bounds(self, i):
        start = self.batch_size * i
        return start, min(self.size, self.batch_size * (i + 1))

    def get_raw_iterator(self):
        self.current_ind = 0

        def _wrapper():
            while self.current_ind < self.num_batch:
                s, e = self._bounds(self.current_ind)
                yield self.xs[s:e], self.ys[s:e]
                self.current_ind += 1
        return _wrapper()

    def get_iterator(self):
        self.current_ind = 0

        def _wrapper():
            while self.current_ind < self.num_batch:
                s, e = self._bounds(self.current_ind)
                x, y = self.xs[s:e], self.ys[s:e]
                yield (x[..., :self.input_dim].contiguous(), y[..., :self.output_dim].contiguous(),
                       y[..., self.output_dim:].contiguous())
                self.current_ind += 1
        return _wrapper()
