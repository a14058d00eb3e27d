"""CPU tests of the dynamic batcher behind ``ASRModel.transcribe`` (no GPU: a fake engine records what it is given).
The contract under test is the one the reference's callers rely on (/root/reference/backend/main.py:429-445: three
executor threads; transcription_manager.py:58: the event loop) — every caller gets exactly its own result, errors reach
the caller that triggered the batch, and a released model raises."""
import threading
import time

import numpy as np
import pytest

from sonicscribe_b200.batcher import DynamicBatcher, Request, bucket_of


class FakeEngine:
    def __init__(self, delay=0.01, fail_on=None):
        self.batches = []
        self.delay = delay
        self.fail_on = fail_on

    def run(self, reqs):
        self.batches.append([(r.wav.shape[0], r.s16, r.max_new) for r in reqs])
        time.sleep(self.delay)
        if self.fail_on is not None and any(r.wav.shape[0] == self.fail_on for r in reqs):
            raise RuntimeError("device fault")
        g = max(r.max_new for r in reqs)
        # "token ids" derived from the segment length so every caller can recognise its own result
        return [[r.wav.shape[0] + t for t in range(g)] for r in reqs], {"decode_ms": 1.0, "batch_size": len(reqs)}


def _req(n, max_new=8, s16=False):
    return Request(np.zeros(n, np.int16 if s16 else np.float32), s16, [1, 2, 3], max_new)


def test_buckets():
    assert bucket_of(15) == 16 and bucket_of(16) == 16 and bucket_of(17) == 64 and bucket_of(150) == 256 and bucket_of(256) == 256


def test_single_caller_is_dispatched_alone_and_immediately():
    eng = FakeEngine(delay=0.0)
    b = DynamicBatcher(eng.run, max_batch=8, window_s=0.5)          # a long window must not delay a lone caller
    t0 = time.perf_counter()
    r = _req(100, 5)
    b.submit_many([r]); b.wait([r])
    assert time.perf_counter() - t0 < 0.2
    assert r.ids == [100, 101, 102, 103, 104] and eng.batches == [[(100, False, 5)]]
    b.close()


def test_concurrent_callers_are_coalesced_and_get_their_own_results():
    eng = FakeEngine(delay=0.02)
    b = DynamicBatcher(eng.run, max_batch=16, window_s=0.01)
    out = {}

    def call(i):
        for rep in range(4):
            r = _req(1000 + i, 6)
            b.submit_many([r]); b.wait([r])
            out[(i, rep)] = r.ids

    th = [threading.Thread(target=call, args=(i,)) for i in range(12)]
    [t.start() for t in th]; [t.join() for t in th]
    for (i, rep), ids in out.items():
        assert ids == [1000 + i + t for t in range(6)]
    st = b.stats()
    assert st["requests"] == 48 and st["max_batch_seen"] >= 6 and st["batches"] < 48
    b.close()


def test_grouping_by_bucket_and_dtype_and_truncation():
    eng = FakeEngine(delay=0.05)
    b = DynamicBatcher(eng.run, max_batch=8, window_s=0.0)
    blocker = _req(1, 4)
    b.submit_many([blocker])                     # keeps the worker busy while the rest queues up
    time.sleep(0.01)
    reqs = [_req(10, 15), _req(11, 150), _req(12, 12), _req(13, 15, s16=True), _req(14, 200)]
    b.submit_many(reqs)
    b.wait([blocker] + reqs)
    assert [len(r.ids) for r in reqs] == [15, 150, 12, 15, 200]      # each keeps the prefix it asked for
    groups = [sorted(x[0] for x in g) for g in eng.batches[1:]]
    assert groups == [[10, 12], [11, 14], [13]]                      # (float, <=16), (float, <=256), (int16, <=16)
    assert reqs[0].ids[:12] == [10 + t for t in range(12)]
    b.close()


def test_max_batch_is_respected_and_order_kept():
    eng = FakeEngine(delay=0.01)
    b = DynamicBatcher(eng.run, max_batch=4, window_s=0.0)
    reqs = [_req(100 + i, 8) for i in range(10)]
    b.submit_many(reqs); b.wait(reqs)
    assert all(len(g) <= 4 for g in eng.batches)
    assert [x[0] for g in eng.batches for x in g] == [100 + i for i in range(10)]
    b.close()


def test_error_reaches_the_callers_of_that_batch_only():
    eng = FakeEngine(delay=0.0, fail_on=666)
    b = DynamicBatcher(eng.run, max_batch=4, window_s=0.0)
    bad = _req(666)
    b.submit_many([bad])
    with pytest.raises(RuntimeError, match="device fault"):
        b.wait([bad])
    ok = _req(5)
    b.submit_many([ok]); b.wait([ok])
    assert ok.ids[0] == 5
    b.close()


def test_closed_batcher_raises():
    b = DynamicBatcher(FakeEngine().run, max_batch=4)
    b.close()
    with pytest.raises(RuntimeError, match="released"):
        b.submit_many([_req(5)])


def test_short_window_requests_never_share_a_batch_with_full_window_requests():
    """The opt-in short encoder window (SONIC_FLAG_SHORT_WINDOW) is a property of the whole device call, so the batcher keys its
    groups on it: interim calls that asked for it are batched together, everything else keeps the reference's full window."""
    seen = []

    def run(reqs):
        seen.append(sorted({r.short for r in reqs}))
        time.sleep(0.02)
        return [[r.wav.shape[0]] * r.max_new for r in reqs], {}

    b = DynamicBatcher(run, max_batch=16, window_s=0.02)
    reqs = [Request(np.zeros(100 + i, np.float32), False, [1], 4, short=(i % 2 == 0)) for i in range(12)]
    th = [threading.Thread(target=lambda r=r: (b.submit_many([r]), b.wait([r]))) for r in reqs]
    [t.start() for t in th]; [t.join() for t in th]
    assert all(len(s) == 1 for s in seen), seen                     # no batch mixed the two kinds
    assert [True] in seen and [False] in seen
    assert all(r.ids == [100 + i] * 4 for i, r in enumerate(reqs))
    assert Request(np.zeros(1, np.float32), False, [1], 4).short is False       # default: the reference's full window
    b.close()
