"""Multi-GPU plumbing of the hot path: reads shard naturally across GPUs (one process per GPU,
replicated reference/index, no collective on the data path -- SURVEY.md 8(e)).  The only
communication is the job-level reduction of the result line: max of the per-rank device time,
sum of the per-rank work counters.  Works over NCCL (GPU tensors) and gloo (CPU tensors)."""
import os


def rank_info():
    return (int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0)))


def shard_seed(base_seed, rank):
    """Every rank simulates its own shard of read pairs (weak scaling): disjoint RNG streams."""
    return int(base_seed) + 1000 * int(rank)


def shard_slice(n_items, rank, world):
    """Contiguous slice of a shared list of seam calls (strong scaling / replay of one dataset)."""
    per = (n_items + world - 1) // world
    lo = min(n_items, rank * per)
    return lo, min(n_items, lo + per)


def reduce_job(times_ms, counters, device=None):
    """times_ms -> elementwise MAX over ranks, counters -> elementwise SUM over ranks."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(times_ms), dtype=torch.float64, device=device)
    c = torch.tensor(list(counters), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return [float(x) for x in t], [float(x) for x in c]
