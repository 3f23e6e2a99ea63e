"""CPU, gloo, world_size 2: the batch-sharded gradient reduction issues ONE collective over the flat buffer that all
parameter gradients alias (the layout mcrn_backward produces), and equals the single-process mean."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from megacrn_b200.ddp import allreduce_gradients, flat_grad_view, shard_batch
    torch.manual_seed(0)
    shapes = [(20, 64), (64, 64), (207, 20), (390, 128), (128,), (1,)]
    params = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    # gradients as views of one flat buffer with 64-float padding, exactly like _MegaCRNFunction.backward
    sizes = [p.numel() for p in params]
    padded = [(s + 63) // 64 * 64 for s in sizes]
    flat = torch.zeros(sum(padded))
    off = 0
    g = torch.Generator().manual_seed(100 + rank)
    for p, s, ps in zip(params, sizes, padded):
        p.grad = flat[off:off + s].view(p.shape)
        p.grad.copy_(torch.randn(p.shape, generator=g))
        off += ps
    assert flat_grad_view(params) is not None
    n = allreduce_gradients(params)
    # second layout: independent gradient tensors -> flatten/copy-back path, still one collective
    params2 = [torch.nn.Parameter(torch.randn(s)) for s in shapes]
    g2 = torch.Generator().manual_seed(200 + rank)
    for p in params2:
        p.grad = torch.randn(p.shape, generator=g2)
    assert flat_grad_view(params2) is None
    n2 = allreduce_gradients(params2)
    x = torch.arange(8.0).reshape(8, 1)
    shard = shard_batch(x, rank, world)
    if rank == 0:
        out.put((n, n2, [p.grad.clone() for p in params], [p.grad.clone() for p in params2], shard.clone()))
    dist.barrier()
    dist.destroy_process_group()


def _run_world(world):
    """One attempt; None when the rendezvous itself failed (the probed port can be taken before rank 0 binds it)."""
    import queue
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        res = q.get(timeout=180)
    except queue.Empty:
        res = None
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.kill()
            res = None
        elif p.exitcode != 0:
            res = None
    return res


def test_single_allreduce_over_flat_gradient_buffer():
    world = 2
    for attempt in range(3):
        res = _run_world(world)
        if res is not None:
            break
    assert res is not None, "gloo world of 2 failed three times"
    n, n2, grads, grads2, shard = res
    assert n == 1 and n2 == 1
    shapes = [g.shape for g in grads]
    for i, s in enumerate(shapes):
        exp = sum(torch.randn(s, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
        # generators advance across tensors: rebuild sequentially
    for seed0, got in ((100, grads), (200, grads2)):
        gens = [torch.Generator().manual_seed(seed0 + r) for r in range(world)]
        for t in got:
            exp = sum(torch.randn(t.shape, generator=gg) for gg in gens) / world
            assert torch.allclose(t, exp, atol=1e-6)
    assert torch.equal(shard, torch.arange(4.0).reshape(4, 1))
