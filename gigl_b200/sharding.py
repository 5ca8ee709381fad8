"""Root sharding for N > 1 GPUs: sampling units (roots) are independent, so rank r of P
owns the contiguous id range [r*N/P, (r+1)*N/P) - the rule the reference uses for its loaders
(python/gigl/distributed/distributed_neighborloader.py:195-216) - and no data-path collective is needed.
The only cross-rank step of a measurement is the max-over-ranks of the device time."""
from __future__ import annotations

import numpy as np


def root_range(n_nodes: int, rank: int, world: int):
    return rank * n_nodes // world, (rank + 1) * n_nodes // world


def root_batches(n_nodes: int, rank: int, world: int, batch: int, n_steps: int, start_step: int = 0):
    """Step s of rank r = the next `batch` ids of the rank's range, in id order, wrapping inside the range."""
    lo, hi = root_range(n_nodes, rank, world)
    span = max(hi - lo, 1)
    return [(lo + (np.arange(batch, dtype=np.int64) + s * batch) % span).astype(np.int32)
            for s in range(start_step, start_step + n_steps)]


def max_over_ranks(value: float, device=None) -> float:
    """The contract's timing reduction: every rank reports the slowest rank's time."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
