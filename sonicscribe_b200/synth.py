"""Deterministic synthetic 16 kHz audio: white noise, speech-like gated harmonics with hard-zero gaps,
and edge cases.  Used by the tests, the golden-fixture generator and bench.py (there is no dataset access)."""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000


def synth_audio(kind: str, n: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / SAMPLE_RATE
    if kind == "noise":
        x = np.clip(rng.standard_normal(n) * 0.1, -1, 1)
    elif kind == "speech":
        x = np.zeros(n)
        for _ in range(int(rng.integers(3, 6))):
            f0 = rng.uniform(100, 400)
            gate = 0.5 * (1 + np.sin(2 * np.pi * rng.uniform(2, 4) * t + rng.uniform(0, 6.28)))
            for h in range(1, 6):
                x += (0.2 / h) * gate * np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 6.28))
        x += 0.01 * rng.standard_normal(n)
        # hard-zero gaps of >= 1 s when there is room
        if n > 3 * SAMPLE_RATE:
            for _ in range(2):
                s = int(rng.integers(0, n - SAMPLE_RATE))
                x[s : s + SAMPLE_RATE] = 0.0
        x = np.clip(x * 0.5, -1, 1)
    elif kind == "zeros":
        x = np.zeros(n)
    elif kind == "impulse":
        x = np.zeros(n)
        x[n // 3] = 0.9
    elif kind == "square":
        x = np.where(np.sin(2 * np.pi * 440.0 * t) >= 0, 1.0, -1.0)
    else:
        raise ValueError(kind)
    return x.astype(np.float32)
