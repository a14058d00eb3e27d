"""Dynamic batching of concurrent ``ASRModel.transcribe`` calls.

The reference serves every segment as its own batch-1 ``generate`` call: up to three executor threads
(/root/reference/backend/main.py:429-445,585-624) plus synchronous calls from the event loop
(/root/reference/backend/transcription_manager.py:58) all enter ``transcribe`` on one model instance.  On a B200 a
batch-1 greedy step streams the whole 2.9 GB decoder for one token; sixteen segments decoded together cost almost the
same.  This module sits behind the UNCHANGED ``transcribe`` signature: callers enqueue a request and block, one worker
thread per replica packs whatever is queued (up to ``max_batch``) into a single ``sonic_transcribe_batch`` call.

Policy
* zero added latency for a lone caller: a request that finds the engine idle and no recent concurrency is dispatched
  immediately as a batch of one;
* while a batch runs, new requests accumulate and form the next batch ("batch whatever is queued");
* when recent concurrency was higher than what is queued, the worker waits up to ``window_s`` (default 3 ms) for the
  stragglers — N looping callers converge to batches of N instead of alternating halves;
* requests are grouped by (sample dtype, max_new_tokens bucket, encoder window): the batch decodes max(max_new_tokens) steps and every
  request keeps the prefix it asked for — greedy decoding is causal, so that prefix is identical to a solo run.
"""
from __future__ import annotations

import collections
import threading
import time
from typing import Callable, List, Optional, Sequence

_BUCKETS = (16, 64, 128, 256, 1 << 30)


def bucket_of(max_new_tokens: int) -> int:
    for b in _BUCKETS:
        if max_new_tokens <= b:
            return b
    return _BUCKETS[-1]


class Request:
    __slots__ = ("wav", "s16", "prompt", "max_new", "short", "done", "ids", "error", "info")

    def __init__(self, wav, s16: bool, prompt: Sequence[int], max_new: int, short: bool = False):
        self.wav, self.s16, self.prompt, self.max_new = wav, bool(s16), prompt, int(max_new)
        self.short = bool(short)          # opt-in short encoder window (interim calls); never mixed with full-window requests
        self.done = threading.Event()
        self.ids: Optional[List[int]] = None
        self.error: Optional[BaseException] = None
        self.info: dict = {}


class DynamicBatcher:
    """``run_batch(requests) -> (list of id lists, info dict)`` is called from the single worker thread."""

    def __init__(self, run_batch: Callable, max_batch: int, window_s: float = 0.003, name: str = "sonic-batcher"):
        self._run = run_batch
        self.max_batch = max(1, int(max_batch))
        self.window_s = max(0.0, float(window_s))
        self._q: collections.deque = collections.deque()
        self._cv = threading.Condition()
        self._closed = False
        self._inflight = 0
        self._recent = collections.deque()         # (time, in-flight count) samples of the last 2 s
        self.batches = 0
        self.requests = 0
        self.max_seen_batch = 0
        self._thread = threading.Thread(target=self._loop, name=name, daemon=True)
        self._thread.start()

    # -- caller side ------------------------------------------------------------------------------------------------
    def submit_many(self, reqs: Sequence[Request]) -> None:
        with self._cv:
            if self._closed:
                raise RuntimeError("ASR model has been released")
            self._q.extend(reqs)
            self._inflight += len(reqs)
            self._recent.append((time.monotonic(), self._inflight))
            self._cv.notify_all()

    def wait(self, reqs: Sequence[Request]) -> None:
        for r in reqs:
            r.done.wait()
        first = next((r.error for r in reqs if r.error is not None), None)
        if first is not None:
            raise first

    # -- worker -----------------------------------------------------------------------------------------------------
    def _recent_peak(self, now: float) -> int:
        while self._recent and now - self._recent[0][0] > 2.0:
            self._recent.popleft()
        return max((c for _, c in self._recent), default=0)

    def _take_group(self) -> List[Request]:
        """Pop the head request and every queued request of the same (dtype, bucket, encoder window), in arrival order."""
        head = self._q[0]
        key = (head.s16, bucket_of(head.max_new), head.short)
        group, rest = [], collections.deque()
        while self._q:
            r = self._q.popleft()
            if len(group) < self.max_batch and (r.s16, bucket_of(r.max_new), r.short) == key:
                group.append(r)
            else:
                rest.append(r)
        self._q = rest
        return group

    def _loop(self):
        while True:
            with self._cv:
                while not self._q and not self._closed:
                    self._cv.wait()
                if self._closed and not self._q:
                    return
                # stragglers: recent concurrency says more callers are about to arrive
                if self.window_s > 0:
                    deadline = time.monotonic() + self.window_s
                    while True:
                        now = time.monotonic()
                        target = min(self._recent_peak(now), self.max_batch)
                        if len(self._q) >= target or now >= deadline or self._closed:
                            break
                        self._cv.wait(deadline - now)
                group = self._take_group()
            try:
                results, info = self._run(group)
                for r, ids in zip(group, results):
                    r.ids = ids[: r.max_new]
                    r.info = dict(info)
            except BaseException as e:       # noqa: BLE001 - delivered to every caller of the batch
                for r in group:
                    r.error = e
            finally:
                with self._cv:
                    self._inflight -= len(group)
                    self.batches += 1
                    self.requests += len(group)
                    self.max_seen_batch = max(self.max_seen_batch, len(group))
                for r in group:
                    r.done.set()

    def close(self):
        with self._cv:
            self._closed = True
            pending = list(self._q)
            self._q.clear()
            self._cv.notify_all()
        for r in pending:
            r.error = RuntimeError("ASR model has been released")
            r.done.set()
        if threading.current_thread() is not self._thread:
            self._thread.join(timeout=30)

    def stats(self) -> dict:
        return {"batches": self.batches, "requests": self.requests, "max_batch_seen": self.max_seen_batch,
                "mean_batch": (self.requests / self.batches) if self.batches else 0.0}
