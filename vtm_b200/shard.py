"""Sharding of independent frame pairs over ranks: pair p belongs to rank p % world.  No data-path
collective exists on this path; the process group is only used for barriers and for the max-over-ranks of timings."""


def pairs_for_rank(n_pairs, world, rank):
    """Indices of the frame pairs rank `rank` of `world` searches."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_pairs, world))


def gather_order(n_pairs, world):
    """Concatenating the per-rank result lists in rank order gives pairs in this order (for re-assembly on rank 0)."""
    return [p for r in range(world) for p in pairs_for_rank(n_pairs, world, r)]


def max_over_ranks(value, device=None):
    """MAX all-reduce of a python float (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
