"""Multi-GPU plumbing for the session-sharded engine.

Sessions are independent, so the only collective on this path is the one-off weight broadcast
at init; afterwards every rank serves its own sessions (sticky placement, weak scaling).
"""
from __future__ import annotations

from typing import Dict, List


def shard_streams(n_streams: int, world: int) -> List[List[int]]:
    """Least-loaded placement of stream ids 0..n-1 over `world` ranks (contiguous, balanced)."""
    base, extra = divmod(n_streams, world)
    out, s = [], 0
    for r in range(world):
        k = base + (1 if r < extra else 0)
        out.append(list(range(s, s + k)))
        s += k
    return out


class StickyPlacement:
    """session_id -> rank, decided once at open time (least loaded), never migrated."""

    def __init__(self, world: int):
        self.load = [0] * world
        self.where: Dict[int, int] = {}

    def open(self, session_id: int) -> int:
        r = min(range(len(self.load)), key=lambda i: (self.load[i], i))
        self.load[r] += 1
        self.where[session_id] = r
        return r

    def close(self, session_id: int) -> None:
        self.load[self.where.pop(session_id)] -= 1


def broadcast_blob(blob, src: int = 0):
    """Broadcast the packed weight arena (a flat uint8 tensor on the rank's device) from `src`.
    NCCL on GPUs; the same call runs over gloo on CPU tensors in tests."""
    import torch.distributed as dist
    dist.broadcast(blob, src=src)
    return blob


def max_over_ranks(value: float, device=None) -> float:
    """Device time is reported as the max over ranks (bench contract)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
