"""2-GPU data parallelism: the batch sharded over two ranks, one NCCL gradient
all-reduce, gives the loss and the 14 gradients of ONE rank on the concatenated batch -- including the masked-MAE
normaliser (model/utils.py:126-133), which is a property of the global batch: the two shards below hold very different
numbers of masked (zero) labels.  Run with `gpurun --gpus 2`; skipped on fewer than two devices."""
import os
import tempfile

import pytest
import torch

from golden_util import rel_l2
from oracle import megacrn_oracle as O

pytestmark = pytest.mark.gpu

SCALER = dict(scaler_mean=50.0, scaler_std=25.0)      # labels == -2.0 map to exactly 0 -> masked
DIMS = dict(num_nodes=60, horizon=4, rnn_units=64, mem_num=8, mem_dim=64)
B, T = 8, 4
FLAGS = [True, False, True, False]


def _data():
    d = O.Dims(**DIMS)
    x, y_cov, labels = O.synthetic_batch(d, B, T, seed=21)
    g = torch.Generator().manual_seed(5)
    drop = torch.rand(labels.shape, generator=g)
    labels[:B // 2][drop[:B // 2] < 0.05] = -2.0      # rank 0: 5 % masked
    labels[B // 2:][drop[B // 2:] < 0.60] = -2.0      # rank 1: 60 % masked
    return d, x, y_cov, labels


def _build(d, dev):
    from megacrn_b200 import MegaCRN
    m = MegaCRN(d.num_nodes, d.input_dim, d.output_dim, d.horizon, d.rnn_units, mem_num=d.mem_num, mem_dim=d.mem_dim).to(dev)
    m.load_state_dict(O.init_params(d, seed=3))
    return m.train()


def _worker(rank, world, port, out_path, graphed):
    import torch.distributed as dist
    from megacrn_b200.ddp import allreduce_gradients, shard_batch
    from megacrn_b200.train_step import GraphedTrainStep, train_step
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    d, x, y_cov, labels = _data()
    m = _build(d, dev)
    xs, ys, ls = (shard_batch(t, rank, world).to(dev) for t in (x, y_cov, labels))
    params = list(m.parameters())
    if graphed:
        g = GraphedTrainStep(m, B // world, T, **SCALER)       # collectives captured inside the graph
        g.load(xs, ys, ls)
        for _ in range(2):
            loss = g(teacher_forcing=FLAGS)
        collectives = 1 if g.allreduce else allreduce_gradients(params)
    else:
        loss = train_step(m, xs, ys, ls, teacher_forcing=FLAGS, **SCALER)
        collectives = allreduce_gradients(params)
    loss = loss.clone()
    dist.all_reduce(loss)
    loss /= world
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({"loss": float(loss.item()), "collectives": collectives,
                    "grads": {n: p.grad.detach().cpu() for n, p in m.named_parameters()}}, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("graphed", [False, True])
def test_two_rank_step_equals_one_rank_on_the_concatenated_batch(graphed):
    import torch.multiprocessing as mp
    from megacrn_b200.train_step import train_step
    d, x, y_cov, labels = _data()
    dev = torch.device("cuda:0")
    m = _build(d, dev)
    loss1 = float(train_step(m, x.to(dev), y_cov.to(dev), labels.to(dev), teacher_forcing=FLAGS, **SCALER).item())
    ref = {n: p.grad.detach().cpu() for n, p in m.named_parameters()}
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "rank0.pt")
        mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, out, graphed), nprocs=2, join=True)
        got = torch.load(out)
    assert got["collectives"] == 1
    assert abs(got["loss"] - loss1) <= 2e-5 * abs(loss1), (got["loss"], loss1)
    for n, g in got["grads"].items():
        assert rel_l2(g, ref[n]) < 2e-4, (n, rel_l2(g, ref[n]))
    # averaging per-rank means instead (the pre-fix behaviour) would be off by the ratio of the mask counts
    cnt = [(labels[:B // 2] != -2.0).sum().item(), (labels[B // 2:] != -2.0).sum().item()]
    assert abs(cnt[0] - cnt[1]) > 0.3 * max(cnt)
