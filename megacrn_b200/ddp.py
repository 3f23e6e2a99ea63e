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
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / world)
        return 1
    buf = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    buf.mul_(1.0 / world)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(buf[off:off + n].view_as(p.grad))
        off += n
    return 1


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous batch shard of rank `rank` (global batch must divide evenly)."""
    b = t.shape[0]
    assert b % world == 0, f"global batch {b} not divisible by world size {world}"
    per = b // world
    return t[rank * per:(rank + 1) * per]
