"""Multi-GPU plumbing for batches of independent worlds.

Worlds never interact (each reference ``World`` owns its bodies, contacts and constraints,
lib/world/world_class.dart:86,93), so a batch shards by contiguous blocks of worlds: one process per GPU,
no collective on the step path.  The only exchange is a fixed-size statistics record per rank, reduced over
NCCL/NVLink (or gloo in the CPU tests) after the timed region.
"""
from __future__ import annotations

from typing import Dict, Tuple


def shard_range(n_units: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block partition: unit u lives on rank u // ceil-balanced block. Returns [begin, end)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, rem = divmod(n_units, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


STAT_KEYS = ("bodies", "body_steps", "contact_iters", "contacts", "rows", "steps")


def reduce_stats(local: Dict[str, float], elapsed_ms: float, device=None) -> Dict[str, float]:
    """Sum the additive statistics over ranks and take the max of the elapsed device time.

    Works without torch.distributed initialised (single process): returns the local values.
    """
    import torch
    import torch.distributed as dist

    out = dict(local)
    out["elapsed_ms"] = float(elapsed_ms)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return out
    t = torch.tensor([float(local.get(k, 0.0)) for k in STAT_KEYS], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    m = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    for k, v in zip(STAT_KEYS, t.tolist()):
        out[k] = v
    out["elapsed_ms"] = float(m.item())
    return out
