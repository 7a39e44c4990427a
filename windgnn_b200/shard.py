"""Sequence-batch sharding across GPUs (one process per GPU, no collective on the data path).

The forward of one window depends on nothing but that window, the shared adjacency and the shared
parameters (``h0 = 0`` for every window: ``src/step6_gcn_gru_combined_model.py:23`` passes no initial
state), so the batch is simply cut into contiguous, near-equal shards.  ``torch.distributed`` is used
only for the barrier and for the max-over-ranks timing of the benchmark.
"""

from __future__ import annotations

from typing import Tuple

import torch


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous ``[lo, hi)`` of ``n`` sequences owned by ``rank`` (sizes differ by at most one)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    if n < 0:
        raise ValueError("n must be >= 0")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """MAX all-reduce of a scalar (identity when no process group is initialised)."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(units_per_rank: int, seconds_local: float, device=None) -> float:
    """Whole-job throughput: units of ALL ranks / the slowest rank's time."""
    import torch.distributed as dist

    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    return world * units_per_rank / max_over_ranks(seconds_local, device)
