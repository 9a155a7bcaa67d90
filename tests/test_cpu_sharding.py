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


# ------------------------------------------------------------------ graph-mode masking of the training iteration
def _train_step_stub(world_is_initialised=False):
    """A TrainStep with nothing CUDA in it: enough to run its host-side tensor logic on the CPU."""
    from xvr_b200.preprocess import XrayTransforms
    from xvr_b200.trainer import TrainStep

    return TrainStep(None, torch.nn.Linear(1, 1), [], {}, XrayTransforms(6), sdd=1020.0, batch_size=7,
                     standardize_global=world_is_initialised)


def test_masked_standardize_equals_indexing_the_kept_samples():
    """Graph mode keeps dropped samples in the batch (static shapes); Standardize's batch-global min/max must then
    run over the kept samples only, so that kept rows equal XrayTransforms(img[keep]) -- what trainer.py:202-207 does."""
    from xvr_b200.preprocess import XrayTransforms

    g = torch.Generator().manual_seed(1)
    x = torch.rand(7, 1, 6, 6, generator=g) * 3
    x[2] += 10.0  # a dropped sample holding the global maximum ...
    x[5] -= 5.0   # ... and one holding the global minimum
    keep = torch.tensor([True, True, False, True, True, False, True])
    step = _train_step_stub()
    got = step._standardize_masked(x, keep)
    assert torch.equal(got[keep], XrayTransforms(6)(x[keep]))
    none = step._standardize_masked(x, torch.zeros(7, dtype=torch.bool))
    assert torch.isfinite(none).all()  # nothing kept: the range falls back to [0, 1], no NaN reaches the weights


def _masked_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from xvr_b200.preprocess import XrayTransforms

        g = torch.Generator().manual_seed(2)
        x = torch.rand(7, 1, 6, 6, generator=g) * 3
        keep = torch.tensor([True, False, True, True, False, True, True])
        x[1] += 9.0
        lo, hi = shard_bounds(7, rank, world)
        step = _train_step_stub(world_is_initialised=True)
        mine = step._standardize_masked(x[lo:hi], keep[lo:hi])
        whole = XrayTransforms(6)(x[keep])  # the unsharded, reference-shaped computation
        idx = torch.nonzero(keep).flatten()
        rows = [(whole[j], mine[i - lo]) for j, i in enumerate(idx.tolist()) if lo <= i < hi]
        out[rank] = all(torch.equal(a, b) for a, b in rows) and len(rows) > 0
    finally:
        dist.destroy_process_group()


def test_masked_standardize_is_batch_global_across_ranks():
    world = 2
    port = 31500 + os.getpid() % 2000
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_masked_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    assert res == {0: True, 1: True}
