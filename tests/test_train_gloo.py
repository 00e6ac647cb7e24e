"""CPU, world_size 2 over gloo: the data-parallel part of the training loss -- the classification normaliser
(`sync_cls_avg_factor`) and `num_total_pos` are averaged over ranks exactly as mmdet's `reduce_mean` does
(petr_head.py:598-613); without a process group the values pass through unchanged."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from petr_b200 import train


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    num_pos = torch.tensor([3.0 if rank == 0 else 8.0])
    cls_avg = torch.tensor([10.5 if rank == 0 else 20.5])
    out = (train._reduce_mean(num_pos), train._reduce_mean(cls_avg))
    # the inputs are not modified in place (they are reused for the per-rank weights)
    assert num_pos.item() == (3.0 if rank == 0 else 8.0)
    ret.put((rank, out[0].item(), out[1].item()))
    dist.barrier()
    dist.destroy_process_group()


def test_loss_normalisers_are_averaged_over_ranks():
    assert train._reduce_mean(torch.tensor([7.0])).item() == 7.0          # no process group: identity
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(ret.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, npos, cavg in got:
        assert abs(npos - 5.5) < 1e-6 and abs(cavg - 15.5) < 1e-6, (rank, npos, cavg)
