"""Pose-batch sharding across one process per GPU.

Every DRR is independent given a replicated volume, so the renderer needs no data-path collective: each rank
renders a contiguous slice of the pose batch.  Collectives appear only around it -- the mean of a loss over
unequal per-rank batches (``keep`` filtering, /root/reference/src/xvr/model/trainer.py:202-204,223), the
batch-global min/max of ``Standardize`` and the max-over-ranks of a timing.
"""

import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard", "global_mean", "global_minmax", "max_over_ranks"]


def shard_bounds(n, rank, world):
    """[lo, hi) of rank's slice when n items are split as evenly as possible (first n % world ranks get one more)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(tensor, rank=None, world=None):
    """This rank's slice of a batch-major tensor."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(tensor.shape[0], rank, world)
    return tensor[lo:hi]


def _active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def global_mean(values):
    """Mean of a per-sample tensor over ALL ranks' samples (per-rank counts may differ)."""
    s = torch.stack([values.sum().double(), torch.tensor(float(values.numel()), dtype=torch.float64, device=values.device)])
    if _active():
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return (s[0] / s[1]).to(values.dtype)


def global_minmax(x):
    """Batch-global (min, max) as Standardize needs them when the batch is sharded."""
    lo, hi = torch.aminmax(x)
    if _active():
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return lo, hi


def max_over_ranks(value, device="cpu"):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if _active():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()
