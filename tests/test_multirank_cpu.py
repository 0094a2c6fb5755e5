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


def test_single_process_is_identity():
    assert shard.max_over_ranks([3.0, 4.0]) == [3.0, 4.0]
    assert shard.local_sessions(0, 1, 4) == [0, 1, 2, 3]
    assert [shard.owner(s, 8) for s in (0, 7, 8, 9)] == [0, 7, 0, 1]
