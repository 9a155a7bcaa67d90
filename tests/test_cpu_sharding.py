"""CPU suite, part 3: the N>1 host logic over the gloo backend, world size 2."""

import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from xvr_b200.sharding import global_mean, global_minmax, max_over_ranks, shard, shard_bounds


def test_shard_bounds_partition_exactly():
    for n in (0, 1, 7, 116, 117, 256):
        for world in (1, 2, 3, 4, 8):
            cuts = [shard_bounds(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1
    assert [hi - lo for lo, hi in (shard_bounds(116, r, 8) for r in range(8))] == [15, 15, 15, 15, 14, 14, 14, 14]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        batch = torch.rand(7, 1, 4, 4, generator=g)  # 7 poses -> ranks hold 4 and 3
        mine = shard(batch)
        lo, hi = global_minmax(mine)
        mean = global_mean(mine.flatten(1).mean(1))
        slow = max_over_ranks(10.0 + rank)
        out[rank] = (mine.shape[0], lo.item(), hi.item(), mean.item(), slow,
                     batch.min().item(), batch.max().item(), batch.flatten(1).mean(1).mean().item())
    finally:
        dist.destroy_process_group()


def test_sharded_reductions_match_the_unsharded_batch():
    world = 2
    port = 29500 + os.getpid() % 2000
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert [res[r][0] for r in range(world)] == [4, 3]
    for r in range(world):
        n, lo, hi, mean, slow, tlo, thi, tmean = res[r]
        assert lo == tlo and hi == thi  # Standardize's batch-global min/max survive sharding bit-exactly
        assert abs(mean - tmean) < 1e-6  # mean over unequal shards == mean over the batch
        assert slow == 11.0
