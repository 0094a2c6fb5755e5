"""Session-shard data parallelism: sessions are independent, so rank r owns sessions
s == r (mod world); every rank holds a full weight replica and there is NO collective on the data path.
torch.distributed is used only to agree on timings (max over ranks) and to barrier the timed region."""
from __future__ import annotations

from typing import Sequence

import torch


def owner(session_id: int, world: int) -> int:
    return session_id % world


def local_sessions(rank: int, world: int, total: int) -> list[int]:
    """Global session ids served by `rank` (sticky for the session's lifetime: its KV caches stay on one GPU)."""
    return [s for s in range(total) if owner(s, world) == rank]


def max_over_ranks(values: Sequence[float], device: str = "cpu") -> list[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return t.tolist()


def whole_job_sessions(world: int, audio_s: float, max_ms_per_step: float) -> float:
    """Headline metric: concurrent real-time sessions of the whole job = world * audio_s / slowest rank's step time."""
    return world * audio_s / (max_ms_per_step / 1e3)
