"""Multi-GPU plumbing: worlds are independent (SimulatorState has no cross-world field,
/root/reference/PhysicsSimulatorCore/SimulatorState.fs:11-19), so the batch shards over ranks as contiguous
world ranges with NO data-path collective.  The only collective is the reduction of the per-batch statistics
block (sums + one max), over NCCL on GPUs (gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, Tuple

SUM_KEYS = ("steps", "body_steps", "pairs_tested", "pairs_colliding", "contacts", "fallback_steps", "nan_worlds",
            "contact_overflow", "reference_exceptions", "fallback_margin_steps", "fallback_capacity_steps",
            "fallback_static_steps", "margin_verified_steps")
MAX_KEYS = ("max_penetration",)


def shard_range(n_worlds: int, rank: int, world_size: int) -> Tuple[int, int]:
    """contiguous partition [g*W/G, (g+1)*W/G)"""
    if world_size < 1 or rank < 0 or rank >= world_size:
        raise ValueError("bad rank/world_size")
    return (rank * n_worlds) // world_size, ((rank + 1) * n_worlds) // world_size


def allreduce_stats(stats: Dict[str, float], device=None) -> Dict[str, float]:
    """sum the counters and max the penetration over all ranks; identity when not distributed"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(stats)
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    s = torch.tensor([int(stats[k]) for k in SUM_KEYS], dtype=torch.int64, device=dev)
    m = torch.tensor([float(stats[k]) for k in MAX_KEYS], dtype=torch.float64, device=dev)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    out = {k: int(v) for k, v in zip(SUM_KEYS, s.tolist())}
    out.update({k: float(v) for k, v in zip(MAX_KEYS, m.tolist())})
    return out
