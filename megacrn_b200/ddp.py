"""Batch-sharded data parallelism for the MegaCRN hot path.

One process per GPU, parameters replicated, the batch split across ranks, and exactly ONE
all-reduce per step over the flat fp32 gradient buffer that ``mcrn_backward`` fills (all 14
``.grad`` tensors alias it).  No activation is ever exchanged: the recurrence is independent
per sequence (model/MegaCRN.py:168-194 has no cross-batch op).
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch
import torch.distributed as dist


def flat_grad_view(params: Iterable[torch.nn.Parameter]) -> Optional[torch.Tensor]:
    """If every .grad aliases one storage (the buffer allocated by the backward), return a 1-D
    view of that whole storage; otherwise None."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return None
    st = grads[0].untyped_storage()
    if any(g.untyped_storage().data_ptr() != st.data_ptr() or g.dtype != torch.float32 for g in grads):
        return None
    n = st.nbytes() // 4
    return torch.empty(0, dtype=torch.float32, device=grads[0].device).set_(st, 0, (n,), (1,))


def _allreduce_mean(t: torch.Tensor, world: int, group) -> None:
    """In-place mean over the group.  NCCL averages inside the collective (ncclAvg: no separate scaling kernel on the
    critical path); other backends (gloo in the CPU tests) sum and scale."""
    if t.is_cuda and dist.get_backend(group) == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.mul_(1.0 / world)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group=None) -> int:
    """Average gradients over the data-parallel group with a single collective.
    Returns the number of collectives issued (1, or 0 outside a process group)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 0
    world = dist.get_world_size(group)
    if world == 1:
        return 0
    params = [p for p in params if p.grad is not None]
    flat = flat_grad_view(params)
    if flat is not None:
        _allreduce_mean(flat, world, group)
        return 1
    buf = torch.cat([p.grad.reshape(-1) for p in params])
    _allreduce_mean(buf, world, group)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(buf[off:off + n].view_as(p.grad))
        off += n
    return 1


def global_mask_count_begin(labels: torch.Tensor, scaler_mean: float, scaler_std: float, group=None):
    """Start the all-reduce of the masked-MAE normaliser (model/utils.py:127-128: ``mask.mean()`` over the GLOBAL batch):
    this rank's count of labels with ``labels * std + mean != 0`` (``mcrn_mask_count``), summed over the group.
    Returns None outside a multi-rank process group, else an opaque handle for ``global_mask_count_end``.  The labels are
    known before the forward, so the 4-byte collective overlaps it."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    from . import _abi
    lib = _abi.load()
    lab = labels.detach().to(torch.float32).contiguous()
    cnt = torch.zeros(1, device=lab.device, dtype=torch.float32)
    with torch.cuda.device(lab.device):
        st = lib.mcrn_mask_count(lab.data_ptr(), lab.numel(), scaler_mean, scaler_std, cnt.data_ptr(),
                                 torch.cuda.current_stream(lab.device).cuda_stream)
    _abi.check(st, "mcrn_mask_count")
    work = dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group, async_op=True)
    return cnt, work, dist.get_world_size(group)


def global_mask_count_into(labels: torch.Tensor, scaler_mean: float, scaler_std: float, out: torch.Tensor, group=None) -> None:
    """Synchronous form: ``out[0]`` (a static 1-element device tensor, e.g. one a captured graph reads) = global count of
    non-masked labels / world size."""
    from . import _abi
    lib = _abi.load()
    lab = labels.detach().to(torch.float32).contiguous()
    out.zero_()
    with torch.cuda.device(lab.device):
        st = lib.mcrn_mask_count(lab.data_ptr(), lab.numel(), scaler_mean, scaler_std, out.data_ptr(),
                                 torch.cuda.current_stream(lab.device).cuda_stream)
    _abi.check(st, "mcrn_mask_count")
    _allreduce_mean(out, dist.get_world_size(group), group)


def global_mask_count_end(pending):
    """Wait for the collective started by ``global_mask_count_begin``; returns the per-rank normaliser
    (global count / world size) as a 1-element device tensor, or None."""
    if pending is None:
        return None
    cnt, work, world = pending
    work.wait()
    cnt.mul_(1.0 / world)
    return cnt


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous batch shard of rank `rank` (global batch must divide evenly)."""
    b = t.shape[0]
    assert b % world == 0, f"global batch {b} not divisible by world size {world}"
    per = b // world
    return t[rank * per:(rank + 1) * per]
