"""Shot-level data parallelism: independent shots are sharded over the ranks (one process per GPU);
the only exchange is the reduction of the failure counters.  Works with any
torch.distributed backend (nccl on GPUs, gloo in the CPU tests)."""


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of `total` shots for `rank` of `world` (sizes differ by at most one)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_counters(counts, device=None):
    """Sum a list/1-D tensor of integer counters over all ranks; returns a list of ints on every rank."""
    import torch
    import torch.distributed as dist
    t = torch.as_tensor(counts, dtype=torch.int64)
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t.cpu().tolist()]


def max_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
