"""CPU tests of the session batcher (speech_to_speech_b200/batcher.py) and of handlers sharing one engine through it:
requests of concurrent sessions are merged into one launch, never across different decoder prompts, results and
exceptions are routed back to the session that asked."""
import threading
import time
from threading import Thread

import numpy as np
import pytest

from speech_to_speech_b200.batcher import SessionBatcher, acquire_shared, release_shared
from test_handlers import make_handler, vad


def test_concurrent_requests_share_a_launch_and_results_are_routed():
    seen = []

    def run(key, items):
        seen.append((key, list(items)))
        time.sleep(0.01)
        return [x * 10 for x in items]

    b = SessionBatcher(run, max_batch=8, max_wait_s=0.25)  # generous window: thread start-up jitter on a loaded box
    try:
        out = {}

        def worker(i):
            out[i] = b.call("k", i)

        ths = [Thread(target=worker, args=(i,)) for i in range(8)]
        [t.start() for t in ths]
        [t.join(5) for t in ths]
        assert out == {i: i * 10 for i in range(8)}
        assert b.items_run == 8 and b.batches_run <= 2 and b.largest_batch >= 4  # 8 arrivals inside the window
        assert all(k == "k" for k, _ in seen)
    finally:
        b.close()


def test_full_batch_is_dispatched_without_waiting_and_overflow_is_requeued():
    sizes = []
    b = SessionBatcher(lambda k, it: (sizes.append(len(it)), list(it))[1], max_batch=4, max_wait_s=5.0)
    try:
        futs = [b.submit("k", i) for i in range(8)]  # two full batches: must not wait for the 5 s window
        t0 = time.monotonic()
        assert [f.result(2) for f in futs] == list(range(8))
        assert time.monotonic() - t0 < 2 and sizes == [4, 4]
    finally:
        b.close()


def test_single_request_is_flushed_after_the_window():
    b = SessionBatcher(lambda k, it: list(it), max_batch=16, max_wait_s=0.02)
    try:
        t0 = time.monotonic()
        assert b.call("k", 7, timeout=2) == 7
        assert 0.015 <= time.monotonic() - t0 < 1.0
    finally:
        b.close()


def test_different_keys_are_never_mixed():
    seen = []
    b = SessionBatcher(lambda k, it: (seen.append((k, tuple(it))), [(k, x) for x in it])[1], max_batch=8, max_wait_s=0.03)
    try:
        futs = [b.submit("en" if i % 2 else "de", i) for i in range(10)]
        res = [f.result(2) for f in futs]
        assert res == [("en" if i % 2 else "de", i) for i in range(10)]
        for k, items in seen:
            assert all((("en" if i % 2 else "de") == k) for i in items)
    finally:
        b.close()


def test_exception_reaches_every_waiter_and_the_batcher_survives():
    calls = {"n": 0}

    def run(key, items):
        calls["n"] += 1
        if calls["n"] == 1:
            raise RuntimeError("device error")
        return list(items)

    b = SessionBatcher(run, max_batch=4, max_wait_s=0.02)
    try:
        futs = [b.submit("k", i) for i in range(3)]
        for f in futs:
            with pytest.raises(RuntimeError, match="device error"):
                f.result(2)
        assert b.call("k", 5, timeout=2) == 5
    finally:
        b.close()


def test_close_serves_the_queue_and_rejects_new_work():
    b = SessionBatcher(lambda k, it: list(it), max_batch=4, max_wait_s=10.0)
    f = b.submit("k", 1)
    b.close()
    assert f.result(1) == 1
    with pytest.raises(RuntimeError):
        b.submit("k", 2)


def test_shared_registry_builds_once_and_closes_on_last_release():
    built, closed = [], []
    a = acquire_shared(("x", 1), lambda: built.append(1) or object(), lambda v: closed.append(v))
    c = acquire_shared(("x", 1), lambda: built.append(1) or object(), lambda v: closed.append(v))
    assert a is c and built == [1]
    release_shared(("x", 1))
    assert closed == []
    release_shared(("x", 1))
    assert closed == [a]


class BatchEngine:
    """Fake engine that records how many utterances each launch carried."""

    def __init__(self):
        self.batches = []
        self.lock = threading.Lock()

    def transcribe(self, audio, opts):
        with self.lock:
            self.batches.append((len(audio), tuple(opts.prefix)))
        time.sleep(0.005)
        return [[len(a), 7, opts.eos_id] for a in audio]

    def detect_language_host(self, audio, sot, lang_ids):
        return lang_ids[0]

    def close(self):
        pass


def test_handlers_of_concurrent_sessions_share_one_engine_launch():
    """6 handler instances (= 6 pipeline units of the reference) on one bundle: their utterances are merged, each
    session gets its own transcription back (the fake encodes the utterance length in the first id)."""
    eng = BatchEngine()
    api, h0 = make_handler("en", max_batch=8, engine=eng)
    handlers = [h0]
    for _ in range(5):
        _, h = make_handler("en", max_batch=8, engine=eng)
        h.bundle.close()              # drop the private bundle the fixture built ...
        h.bundle = h0.bundle          # ... and share the first one, as acquire_shared() does in setup()
        handlers.append(h)
    outs = {}

    def session(i):
        outs[i] = list(handlers[i].process(vad(api, n=16000 + 160 * i)))[0]

    ths = [Thread(target=session, args=(i,)) for i in range(6)]
    [t.start() for t in ths]
    [t.join(5) for t in ths]
    for i in range(6):
        assert outs[i].text.split()[0] == str(16000 + 160 * i)
    assert sum(n for n, _ in eng.batches) == 6 and max(n for n, _ in eng.batches) >= 3
    assert len({p for _, p in eng.batches}) == 1
    h0.bundle.close()


def test_llm_decode_chunks_of_concurrent_sessions_share_a_launch():
    """3 sessions streaming from ONE shared Llama bundle: their decode chunks are merged into multi-session launches
    (the fake engine records the slots per launch) and every session still gets its own token sequence."""
    import torch
    from types import SimpleNamespace
    from speech_to_speech_b200.handlers.language_model_handler import TokenStreamer, _IdTokenizer, _LlamaBundle

    class FakeLlama:
        device = "cpu"
        cfg = SimpleNamespace(max_prefill=64)

        def __init__(self):
            self.launches, self.state = [], {}

        def max_decode_batch(self):
            return 4

        def reset(self, slot):
            self.state[slot] = 0

        def prefill(self, slot, ids):
            self.state[slot] = 1000 * (slot + 1)          # session s generates 1000(s+1), 1000(s+1)+1, ...
            return torch.tensor([self.state[slot]]), None

        def decode(self, slots, first, n, eos_id=-1):
            self.launches.append(list(slots))
            time.sleep(0.003)
            rows = []
            for s, f in zip(slots, first.tolist()):
                rows.append([f + 1 + i for i in range(n)])
            return torch.tensor(rows), torch.tensor([n] * len(slots))

        def close(self):
            pass

    eng = FakeLlama()
    tok = _IdTokenizer(100000)
    real_tensor = torch.tensor
    torch.tensor = lambda data, dtype=None, device=None: real_tensor(data, dtype=dtype)  # no CUDA on the CPU box
    try:
        bundle = _LlamaBundle(eng, tok, [7], max_sessions=3, batch_wait_s=0.2)
        outs = {}

        def session(i):
            slot = bundle.acquire_slot()
            st = TokenStreamer(eng, lambda ids: tok.decode(ids), [7], chunk=4, slot=slot, decode_chunk=bundle.decode_chunk, lock=bundle.lock)
            "".join(st.stream([1, 2, 3], max_new_tokens=13))
            outs[slot] = st.generated
            bundle.release_slot(slot)

        ths = [Thread(target=session, args=(i,)) for i in range(3)]
        [t.start() for t in ths]
        [t.join(10) for t in ths]
    finally:
        torch.tensor = real_tensor
        bundle.close()
    assert sorted(outs) == [0, 1, 2]
    for slot, gen in outs.items():
        assert gen == [1000 * (slot + 1) + i for i in range(13)]
    assert max(len(l) for l in eng.launches) >= 2                 # chunks of different sessions rode one launch
    assert all(len(set(l)) == len(l) for l in eng.launches)       # a session appears at most once per launch


def test_random_arrivals_keep_every_result_with_its_request():
    """Randomised load: 6 producer threads, 3 keys, random gaps; every future gets exactly its own (key, item) back,
    batches never exceed max_batch and never mix keys."""
    import random
    seen = []

    def run(key, items):
        seen.append((key, len(items)))
        time.sleep(random.random() * 0.002)
        return [(key, it) for it in items]

    b = SessionBatcher(run, max_batch=5, max_wait_s=0.003)
    results, lock = {}, threading.Lock()

    def producer(pid):
        rnd = random.Random(pid)
        for n in range(40):
            key = rnd.choice(["en", "de", "fr"])
            item = (pid, n)
            fut = b.submit(key, item)
            if rnd.random() < 0.5:
                time.sleep(rnd.random() * 0.002)
            with lock:
                results[item] = (key, fut)

    try:
        ths = [Thread(target=producer, args=(p,)) for p in range(6)]
        [t.start() for t in ths]
        [t.join(30) for t in ths]
        assert len(results) == 240
        for item, (key, fut) in results.items():
            assert fut.result(5) == (key, item)
        assert sum(n for _, n in seen) == 240 and max(n for _, n in seen) <= 5
        assert b.items_run == 240
    finally:
        b.close()


def test_thread_context_is_entered_by_the_engine_thread_and_failures_reach_the_callers():
    """The lane mechanism: the batcher's engine thread lives inside a context (the lane's CUDA stream in production); a context
    that cannot be entered must fail the requests, not strand the handler threads."""
    import contextlib
    import threading
    from speech_to_speech_b200.batcher import SessionBatcher
    seen = []

    @contextlib.contextmanager
    def ctx():
        seen.append(threading.current_thread().name)
        yield

    b = SessionBatcher(lambda key, items: [threading.current_thread().name for _ in items], 4, 0.001, "lane-thread", thread_context=ctx)
    assert b.call("k", 1, timeout=5) == "lane-thread" and seen == ["lane-thread"]
    b.close()

    def broken():
        raise AttributeError("no device")
    b2 = SessionBatcher(lambda key, items: items, 4, 0.001, "lane-thread-2", thread_context=broken)
    with pytest.raises(RuntimeError, match="could not enter"):
        b2.call("k", 1, timeout=5)
    b2.close()


def test_idle_gap_dispatches_when_the_burst_is_over_not_after_the_whole_window():
    """With idle_gap_s the window is an upper bound: a lone request leaves after one gap, a burst leaves together."""
    import threading
    import time
    from speech_to_speech_b200.batcher import SessionBatcher
    sizes = []
    b = SessionBatcher(lambda k, it: (sizes.append(len(it)), list(it))[1], max_batch=16, max_wait_s=2.0, idle_gap_s=0.02)
    t0 = time.monotonic()
    assert b.call("k", 1, timeout=5) == 1
    assert time.monotonic() - t0 < 1.0 and sizes == [1]          # did not sit out the 2 s window
    futs = []

    def burst():
        for i in range(6):
            futs.append(b.submit("k", i))
            time.sleep(0.002)                                     # arrivals closer together than the gap: one batch
    th = threading.Thread(target=burst)
    th.start()
    th.join()
    assert [f.result(5) for f in futs] == list(range(6)) and sizes == [1, 6]
    b.close()


def test_two_groups_merge_instead_of_alternating_half_full_launches():
    """Two groups of sessions out of phase by one launch: the group that queued up behind the running launch must wait one idle
    gap after it ends, so that the sessions of that launch (which resubmit immediately) ride along -- launches fill up to
    max_batch instead of alternating between the groups forever."""
    import threading
    import time
    from speech_to_speech_b200.batcher import SessionBatcher
    sizes = []

    def run(key, items):
        sizes.append(len(items))
        time.sleep(0.03)                       # a launch
        return list(items)
    b = SessionBatcher(run, max_batch=8, max_wait_s=0.5, idle_gap_s=0.01)

    def session(i, delay):
        time.sleep(delay)
        for _ in range(6):
            b.call("k", i, timeout=10)
    ths = [threading.Thread(target=session, args=(i, 0.0 if i < 4 else 0.015)) for i in range(8)]   # group B starts mid-launch
    [t.start() for t in ths]
    [t.join() for t in ths]
    b.close()
    assert sizes[0] == 4 and sizes.count(8) >= 4, sizes    # after the first launches the groups travel together


def test_units_join_the_lanes_round_robin_when_the_caller_does_not_pin_one():
    from speech_to_speech_b200.batcher import assign_lane
    g = ("kind", "model-%d" % id(object()), 0)
    assert [assign_lane(g, 2) for _ in range(5)] == [0, 1, 0, 1, 0]
    assert [assign_lane(("other",) + g, 3) for _ in range(4)] == [0, 1, 2, 0]      # counters are per (kind, model, device)
    assert assign_lane(g, 1) == 0 and assign_lane(g, 2) == 1                       # one lane does not advance the counter
