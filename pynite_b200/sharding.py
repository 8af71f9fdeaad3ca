"""Load-combination sharding across GPUs: combos are independent given K, so every rank
assembles and factorises K itself and solves its own block of combos; there is no collective on the data
path.  `torch.distributed` is used only to hand every rank the results of the others (all-gather of the
displacement / reaction rows) and for barriers / max-over-ranks timing.  One process per GPU, NCCL on GPUs,
gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def world():
    """(rank, world_size) of the default process group, (0, 1) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def shard_combos(n_combo, rank, world_size):
    """Contiguous, balanced block of combo indices owned by `rank` (first `n_combo % world_size` ranks get one more)."""
    base, extra = divmod(n_combo, world_size)
    start = rank*base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def all_gather_rows(local, counts, device=None):
    """All-gathers row blocks of different heights: `local` is (counts[rank], n); returns (sum(counts), n) in rank
    order on every rank.  Rows are padded to the largest block so one fixed-size all_gather is enough."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return np.asarray(local)
    n = local.shape[1] if local.ndim == 2 else 0
    hmax = max(counts)
    buf = torch.zeros((hmax, n), dtype=torch.float64, device=device)
    if counts[rank]:
        buf[:counts[rank]] = torch.as_tensor(np.ascontiguousarray(local), dtype=torch.float64, device=device)
    out = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(out, buf)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)], axis=0)


def max_over_ranks(x, device=None):
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def any_rank_failed(failed, device=None):
    """Collective OR of a per-rank failure flag: every rank learns whether any rank raised, so that an exception on one
    rank (singular matrix, unstable node) is raised everywhere instead of leaving the others blocked in a collective."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return bool(failed)
    t = torch.tensor([1.0 if failed else 0.0], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(t.item() > 0)
