"""Host-side mirrors of the reference's manager classes (no GPU): thresholds, token budgets and error behaviour of
backend/transcription_manager.py:16-65 and backend/models_manager.py:16-32,52-62."""
import asyncio

import numpy as np
import pytest

import sonicscribe_b200.models_manager as mm
import sonicscribe_b200.transcription_manager as tm
from sonicscribe_b200.config import AppConfig


class FakeASR:
    def __init__(self, fail=False):
        self.calls = []
        self.fail = fail

    def transcribe(self, audio_tensor, sampling_rate=16000, max_new_tokens=128, **kw):
        if self.fail:
            raise RuntimeError("boom")
        self.calls.append((tuple(audio_tensor.shape), str(audio_tensor.dtype), sampling_rate, max_new_tokens,
                           float(audio_tensor.abs().max())))
        return "  hello world \n"

    def close(self):
        pass


@pytest.fixture()
def fake(monkeypatch):
    f = FakeASR()
    monkeypatch.setattr(mm, "_asr_model", f)
    return f


def _pcm(n, amp=16384):
    return (np.full(n, amp, dtype=np.int16)).tobytes()


def test_uninitialised_model_raises_reference_message(monkeypatch):
    monkeypatch.setattr(mm, "_asr_model", None)
    with pytest.raises(RuntimeError, match=r"ASR model not initialized\. Call asr_model_init\(\) first!"):
        mm.asr_model_get()


def test_init_is_idempotent(monkeypatch):
    sentinel = FakeASR()
    monkeypatch.setattr(mm, "_asr_model", sentinel)
    mm.asr_model_init()                                   # must not construct a second model (models_manager.py:26-28)
    assert mm.asr_model_get() is sentinel


def test_temporary_threshold_and_token_budget(fake):
    mgr = tm.TranscriptionManager()
    short = b"\0" * (AppConfig.AUDIO_CHUNK_SIZE - 1)
    assert asyncio.run(mgr.transcribe_temporary(short)) == ""
    assert asyncio.run(mgr.transcribe_temporary(b"")) == ""
    assert fake.calls == []
    n = AppConfig.AUDIO_CHUNK_SIZE            # bytes; int16 -> n / 2 samples
    assert asyncio.run(mgr.transcribe_temporary(_pcm(n // 2))) == "hello world"
    shape, dtype, sr, budget, peak = fake.calls[-1]
    assert shape == (1, n // 2) and dtype == "torch.float32" and sr == 16000 and budget == 15
    assert abs(peak - 0.5) < 1e-6                           # int16 / 32768 (transcription_manager.py:50)


@pytest.mark.parametrize("dur,budget", [(0.0, 50), (1.0, 55), (9.9, 99), (30.0, 200), (100.0, 200)])
def test_committed_token_budget(fake, dur, budget):
    mgr = tm.TranscriptionManager()
    assert asyncio.run(mgr.transcribe_committed(_pcm(AppConfig.AUDIO_CHUNK_SIZE), dur)) == "hello world"
    assert fake.calls[-1][3] == budget                      # min(50 + int(5 * duration), 200)


def test_committed_needs_two_chunks(fake):
    mgr = tm.TranscriptionManager()
    assert asyncio.run(mgr.transcribe_committed(b"\0" * (2 * AppConfig.AUDIO_CHUNK_SIZE - 2), 1.0)) == ""
    assert fake.calls == []


def test_errors_degrade_to_empty_string(monkeypatch):
    monkeypatch.setattr(mm, "_asr_model", FakeASR(fail=True))
    mgr = tm.TranscriptionManager()
    assert asyncio.run(mgr.transcribe_temporary(_pcm(AppConfig.AUDIO_CHUNK_SIZE))) == ""
    assert asyncio.run(mgr.transcribe_committed(_pcm(AppConfig.AUDIO_CHUNK_SIZE * 2), 2.0)) == ""
