"""Sharding of independent segments over model replicas (one replica per GPU, no collective on the data path).

* ``cut_long_segments``  — the fixed-length cutting rule of /root/reference/backend/main.py:527-567 (sample arithmetic only).
* ``shard_indices`` / ``gather_in_order`` — torchrun-style data parallelism: rank r takes segments r, r+W, ...; results are
  re-assembled in segment order (the NDJSON stream of main.py:448-468 is in segment order).
* ``ReplicaPool``        — one process, one host thread + one ``ASRModel`` per visible GPU (ctypes releases the GIL, so the
  replicas run concurrently); replaces the ``Semaphore(3)`` + executor fan-out of main.py:429-445.
"""
from __future__ import annotations

import math
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Callable, List, Sequence


def cut_long_segments(start_sample: int, end_sample: int, sample_rate: int, max_segment_duration: float, min_duration: float = 0.1):
    """[(start, end)] sample ranges: segments no longer than max_segment_duration, pieces <= min_duration dropped
    (main.py:543-565)."""
    duration = (end_sample - start_sample) / sample_rate
    if duration <= max_segment_duration:
        return [(start_sample, end_sample)]
    n = int(math.ceil(duration / max_segment_duration))
    per = int(max_segment_duration * sample_rate)
    out = []
    for i in range(n):
        s = start_sample + i * per
        e = min(start_sample + (i + 1) * per, end_sample)
        if (e - s) / sample_rate > min_duration:
            out.append((s, e))
    return out


def shard_indices(n_items: int, world_size: int, rank: int) -> List[int]:
    return list(range(rank, n_items, world_size))


def gather_in_order(local_results: Sequence, n_items: int, world_size: int, rank: int, group=None) -> List:
    """all_gather the per-rank results (python objects) and restore item order.  Works on gloo (CPU tests) and nccl."""
    import torch.distributed as dist

    if world_size == 1:
        return list(local_results)
    bucket = [None] * world_size
    dist.all_gather_object(bucket, list(local_results), group=group)
    out = [None] * n_items
    for r, res in enumerate(bucket):
        for j, idx in enumerate(shard_indices(n_items, world_size, r)):
            out[idx] = res[j]
    return out


class ReplicaPool:
    """N replicas in one process.  ``factory(device_index)`` builds a model exposing ``transcribe_batch(list, **kw)``."""

    def __init__(self, factory: Callable[[int], object], n_gpus: int, batch: int = 8):
        self.replicas = [factory(i) for i in range(n_gpus)]
        self.batch = batch
        self._locks = [threading.Lock() for _ in self.replicas]
        self._exec = ThreadPoolExecutor(max_workers=n_gpus)

    def transcribe_segments(self, segments: Sequence, **kw) -> List[str]:
        n = len(self.replicas)
        chunks = [(i, segments[i:i + self.batch]) for i in range(0, len(segments), self.batch)]

        def work(k):
            start, chunk = chunks[k]
            g = k % n
            with self._locks[g]:
                return start, self.replicas[g].transcribe_batch(list(chunk), **kw)

        out: List = [None] * len(segments)
        for start, res in self._exec.map(work, range(len(chunks))):
            out[start:start + len(res)] = res
        return out

    def close(self):
        self._exec.shutdown(wait=True)
        for r in self.replicas:
            if hasattr(r, "close"):
                r.close()
