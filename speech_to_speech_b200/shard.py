"""Session-shard data parallelism: sessions are independent, so rank r owns sessions
s == r (mod world) and every rank holds a full weight replica.  The only exchange on the data path is the session-shard
split itself when ingest is centralised (one rank terminates the clients' connections, the reference's single-server
shape): `scatter_from_ingest` hands each rank its sessions' PCM and `gather_to_ingest` brings the first audio blocks / ids
back -- NCCL send/recv over NVLink under torch.distributed's scatter / gather (gloo on the CPU tests).  Payloads are
<= 0.64 MB per session-turn in and a few KB out: latency-bound, never bandwidth-bound.  Timings are agreed with a
max-over-ranks all-reduce."""
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


def _dist_on() -> bool:
    return torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1


def shard_layout(total: int, world: int) -> list[list[int]]:
    """[rank] -> the global session ids it owns, in the order the ingest rank packs them."""
    return [local_sessions(r, world, total) for r in range(world)]


def scatter_from_ingest(full: "torch.Tensor | None", per_rank: int, row_shape: Sequence[int], dtype: torch.dtype, device: str,
                        src: int = 0) -> torch.Tensor:
    """Ingest rank `src` holds `full` [world * per_rank, *row_shape] packed rank-major (shard_layout order); every rank gets
    its [per_rank, *row_shape] block.  One grouped send/recv: the session-shard split of SURVEY.md 8(e)."""
    out = torch.empty((per_rank, *row_shape), dtype=dtype, device=device)
    if not _dist_on():
        out.copy_(full[:per_rank])
        return out
    rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
    chunks = [full[r * per_rank:(r + 1) * per_rank].contiguous() for r in range(world)] if rank == src else None
    torch.distributed.scatter(out, scatter_list=chunks, src=src)
    return out


def gather_to_ingest(local: torch.Tensor, dst: int = 0) -> "torch.Tensor | None":
    """Every rank's [per_rank, ...] result block -> [world * per_rank, ...] on the ingest rank (None elsewhere)."""
    if not _dist_on():
        return local.clone()
    rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
    bufs = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
    torch.distributed.gather(local.contiguous(), gather_list=bufs, dst=dst)
    return torch.cat(bufs, 0) if rank == dst else None
