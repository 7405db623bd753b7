"""Multi-GPU plumbing for the dense-head path: images are independent, so the batch is sharded by image
across ranks (one process per GPU) and the only exchange is one all-reduce of the loss scalars
{cls, reg, cen, n_pos} per step.  Works with any torch.distributed backend
(NCCL over NVLink on the B200 box, gloo in the CPU tests)."""
import numpy as np
import torch


def shard_range(n_images, rank, world_size):
    """Contiguous, balanced slice [lo, hi) of the batch owned by `rank` (earlier ranks take the remainder)."""
    base, rem = divmod(int(n_images), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(arrays, rank, world_size):
    """Slice every [B, ...] array / tensor of `arrays` down to this rank's images."""
    n = len(arrays[0])
    lo, hi = shard_range(n, rank, world_size)
    return [a[lo:hi] for a in arrays]


def allreduce_losses(total, group=None):
    """Sum the per-rank loss vector across ranks in place (no-op when torch.distributed is not initialised)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return total


def gather_counts(per_rank_count, group=None):
    """All ranks learn how many images every rank processed (ragged last shard)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [int(per_rank_count)]
    t = torch.tensor([int(per_rank_count)], dtype=torch.int64, device="cuda" if dist.get_backend(group) == "nccl" else "cpu")
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return [int(x[0]) for x in out]
