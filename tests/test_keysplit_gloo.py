"""CPU, world_size 2 over gloo: the host-side protocol of the key-split over camera views
 -- view sharding, img_metas slicing, the packed [O | m, l] partial layout
and the all-gather + merge math.  The CUDA kernels that produce / merge the partials are checked
on the GPU (tests/test_gpu_kernels.py::test_cross_attention_bf16_masks_and_merge,
tests/test_gpu_head.py::test_keysplit_over_views_matches_single_gpu); here the merge is evaluated
with plain torch on CPU to validate the exchange layout end to end across two processes."""
import math
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from petr_b200 import synth
from petr_b200.configs import WORKLOADS
from petr_b200.engine_tc import shard_views, slice_img_metas


def test_shard_views_partition():
    for n in (6, 12, 7):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                sl = shard_views(n, world, r)
                got += list(range(n))[sl]
            assert got == list(range(n))
    metas = synth.make_img_metas(WORKLOADS["v2-vov-p4-800x320"], 2, 0)
    sub = slice_img_metas(metas, shard_views(12, 4, 1))
    assert len(sub[0]["lidar2img"]) == 3 and sub[0]["lidar2img"][0] is metas[0]["lidar2img"][3]
    assert sub[1]["timestamp"] == metas[1]["timestamp"] and len(sub[1]["img_shape"]) == 3


def _partial(q, k, v):
    """(O un-normalised, m, l) of softmax attention over a key shard, log2 domain, packed like the
    CUDA kernel: [rows*32 | rows*2]."""
    s = (q @ k.transpose(-1, -2)) * math.log2(math.e)
    m = s.max(-1, keepdim=True).values
    p = torch.exp2(s - m)
    o = p @ v
    rows = o.shape[0] * o.shape[1] * o.shape[2]
    return torch.cat([o.reshape(rows * 32), torch.stack([m.reshape(rows), p.sum(-1).reshape(rows)], 1).reshape(-1)])


def _merge(parts, nparts, B, H, Lq):
    rows = B * H * Lq
    parts = parts.view(nparts, rows * 34)
    O = parts[:, :rows * 32].reshape(nparts, rows, 32)
    ml = parts[:, rows * 32:].reshape(nparts, rows, 2)
    m = ml[..., 0].max(0).values
    w = torch.exp2(ml[..., 0] - m)
    return (w[..., None] * O).sum(0) / (w * ml[..., 1]).sum(0)[:, None]


def _worker(rank, world, port, q, k, v, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_views, per_view = 6, k.shape[2] // 6
    sl = shard_views(n_views, world, rank)
    ks = slice(sl.start * per_view, sl.stop * per_view)
    local = _partial(q, k[:, :, ks], v[:, :, ks])
    gathered = torch.empty(world * local.numel())
    dist.all_gather_into_tensor(gathered, local)
    B, H, Lq, _ = q.shape
    out = _merge(gathered, world, B, H, Lq)
    if rank == 0:
        ret.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_keysplit_exchange_two_ranks_gloo():
    torch.manual_seed(0)
    B, H, Lq, Lk = 1, 8, 50, 6 * 20
    q, k, v = torch.randn(B, H, Lq, 32), torch.randn(B, H, Lk, 32), torch.randn(B, H, Lk, 32)
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, k, v, ret)) for r in range(2)]
    for p in procs:
        p.start()
    out = ret.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = (torch.softmax(q @ k.transpose(-1, -2), -1) @ v).reshape(B * H * Lq, 32)
    assert (out - ref).abs().max() <= 1e-5
