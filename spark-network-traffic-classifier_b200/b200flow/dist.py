"""Row sharding over ranks (one process per GPU, torch.distributed): SURVEY.md §8(e).

Rows are independent units; rank r holds a contiguous block of the global row range.  The only
data-path collective of the trainer is the per-level histogram all-reduce (forest.fit_forest); the fit-time
one-offs (category counts, moments, the findSplits sample, confusion counts) are tiny all-reduces/all-gathers.
Works with the NCCL backend on GPUs and with gloo on CPU tensors (used by the CPU tests)."""
import torch
import torch.distributed as dist


def group():
    """the default process group when running under torchrun with world_size > 1, else None."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.group.WORLD
    return None


def shard_bounds(n_total, rank, world):
    """contiguous block partition: rank r owns rows [r*n/G, (r+1)*n/G)."""
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def global_offset(n_local, device, grp=None):
    """(first global row index of this rank's block, global row count) from the local counts."""
    grp = grp if grp is not None else group()
    if grp is None:
        return 0, int(n_local)
    world, rank = dist.get_world_size(grp), dist.get_rank(grp)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([int(n_local)], dtype=torch.int64, device=device), group=grp)
    counts = [int(c.item()) for c in counts]
    return sum(counts[:rank]), sum(counts)


def all_reduce_sum_(t, grp=None):
    """in-place sum over ranks (integer tensors stay exact -> bit-identical models for any world size)."""
    grp = grp if grp is not None else group()
    if grp is not None:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=grp)
    return t
