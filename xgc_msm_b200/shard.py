"""Replicate sharding across GPUs: replicates are independent (the reference runs them as unrelated SLURM
array tasks, burnin/cal/sim1/sim1_batch1.sh:2-9), so rank k owns a contiguous block of global replicate ids and the only
exchange is the final gather of the [replicates x targets] matrix.  Philox is keyed by the GLOBAL replicate id
(xgc_config.replicate_offset), so a replicate's trajectory does not depend on the number of GPUs."""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """(first global replicate id, count) of `rank`; blocks differ by at most one replicate."""
    base, rem = divmod(n_total, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def gather_targets(local: np.ndarray, n_total: int, device=None) -> np.ndarray:
    """All-gather the per-rank [count, K] target matrices into [n_total, K] (torch.distributed: nccl on GPUs, gloo in
    the CPU tests).  Uneven shards are padded to the largest block."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    kmax = max(shard_range(n_total, r, world)[1] for r in range(world))
    t = torch.zeros((kmax, local.shape[1]), dtype=torch.float64, device=device)
    t[: local.shape[0]] = torch.from_numpy(np.ascontiguousarray(local)).to(t.device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    return np.concatenate([parts[r][: shard_range(n_total, r, world)[1]].cpu().numpy() for r in range(world)], axis=0)
