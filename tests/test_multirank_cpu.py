"""CPU, world_size 2, gloo: the N>1 host logic of bench.py (session sharding, barrier, max-over-ranks timing)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from speech_to_speech_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.local_sessions(rank, world, 11)
    dist.barrier()
    ms = 10.0 + 5.0 * rank  # rank 1 is the slow one
    red = shard.max_over_ranks([ms, float(len(mine))])
    q.put((rank, mine, red))
    dist.barrier()
    dist.destroy_process_group()


def test_sessions_partition_and_timings_reduce_to_max():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, red0), (r1, s1, red1) = res
    assert sorted(s0 + s1) == list(range(11)) and not set(s0) & set(s1)  # every session owned exactly once
    assert s0 == [0, 2, 4, 6, 8, 10] and s1 == [1, 3, 5, 7, 9]
    assert red0 == red1 == [15.0, 6.0]  # max over ranks, identical on every rank
    assert shard.whole_job_sessions(2, 10.0, red0[0]) == pytest.approx(2 * 10.0 / 0.015)


def _exchange_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per_rank, n = 3, 5
    layout = shard.shard_layout(world * per_rank, world)
    full = None
    if rank == 0:   # the ingest rank packs session s's row (value = s) rank-major
        full = torch.stack([torch.full((n,), float(s)) for r in range(world) for s in layout[r]])
    mine = shard.scatter_from_ingest(full, per_rank, (n,), torch.float32, "cpu")
    result = (mine[:, :2] * 10).to(torch.int32)                  # "ids" computed by the owner
    back = shard.gather_to_ingest(result)
    q.put((rank, mine[:, 0].tolist(), None if back is None else back[:, 0].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_session_shard_scatter_and_gather_through_the_ingest_rank():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, got0, back0), (r1, got1, back1) = res
    assert got0 == [0.0, 2.0, 4.0] and got1 == [1.0, 3.0, 5.0]           # every rank received exactly the sessions it owns
    assert back1 is None and back0 == [0, 20, 40, 10, 30, 50]            # results return rank-major to the ingest rank


def test_single_process_is_identity():
    full = torch.arange(6, dtype=torch.float32).reshape(3, 2)
    assert torch.equal(shard.scatter_from_ingest(full, 3, (2,), torch.float32, "cpu"), full)
    assert torch.equal(shard.gather_to_ingest(full), full)
    assert shard.max_over_ranks([3.0, 4.0]) == [3.0, 4.0]
    assert shard.local_sessions(0, 1, 4) == [0, 1, 2, 3]
    assert [shard.owner(s, 8) for s in (0, 7, 8, 9)] == [0, 7, 0, 1]
