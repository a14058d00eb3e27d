"""GPU tests of the drop-in class itself (run with -m gpu on a B200): ``sonicscribe_b200.asr.ASRModel`` constructed and
called the way the reference server calls /root/reference/backend/asr.py — ``transcribe`` from several threads at once
(main.py:429-445), the ``return_debug_info`` keys (asr.py:445-465), 1-D / multi-channel input (asr.py:248-252), releasing
``.model`` (main.py:84-88), the manager singletons, a checkpoint DIRECTORY with tokenizer + chat template (asr.py:66-82,
393-399, 425-429) and the replica pool.  Nothing here reads /root/reference."""
import asyncio
import os
import threading

import numpy as np
import pytest
import torch

from oracle import mel_oracle as mo
from oracle import model_oracle as ora
from sonicscribe_b200.asr import ASRModel
from sonicscribe_b200.engine import num_audio_tokens
from sonicscribe_b200.prompt import synthetic_prompt_ids
from sonicscribe_b200.weights import ModelDims, synthetic_state_dict

pytestmark = pytest.mark.gpu

SPEC = "synthetic:seed=0,enc=2,dec=2"


def _segments(k, base=0):
    lens = [163840, 20480, 320000, 48000, 96000, 240000, 32000, 280000]
    return [torch.from_numpy(mo.synth_audio("speech" if i % 2 else "noise", lens[i % len(lens)], seed=base + i))[None] for i in range(k)]


@pytest.fixture(scope="module")
def asr():
    m = ASRModel(SPEC, device="cuda", mode="native", max_batch=8, max_prompt=320, max_new_tokens=64)
    yield m
    m.close()


def _ids(model, x, **kw):
    return model.transcribe(x, return_debug_info=True, **kw)["token_ids"]


def test_transcribe_returns_text_and_debug_info(asr):
    x = _segments(1)[0]
    text = asr.transcribe(x, sampling_rate=16000, max_new_tokens=12)
    assert isinstance(text, str) and text == text.strip() and text.startswith("<")
    info = asr.transcribe(x, max_new_tokens=12, return_debug_info=True)
    for key in ("transcript", "processing_time", "audio_length_sec", "mode", "device", "gpu_memory_allocated_mb", "gpu_memory_reserved_mb"):
        assert key in info, key                                  # the reference's keys (asr.py:445-465)
    assert info["transcript"] == text and info["mode"] == "native" and info["device"].startswith("cuda")
    assert abs(info["audio_length_sec"] - x.shape[-1] / 16000) < 1e-9 and info["processing_time"] > 0
    assert len(info["token_ids"]) == 12 and info["gpu_memory_allocated_mb"] > 100
    gi = asr.get_model_info()
    assert gi["mode"] == "native" and gi["target_sampling_rate"] == 16000 and gi["is_glm_asr"] and "gpu_name" in gi


def test_input_shapes_first_channel_and_numpy(asr):
    x = _segments(1)[0]                                         # [1, N]
    ref = _ids(asr, x, max_new_tokens=10)
    assert _ids(asr, x[0], max_new_tokens=10) == ref            # [N]
    stereo = torch.cat([x, torch.flip(x, dims=[1])], dim=0)     # [2, N]: only the first channel is used (asr.py:252)
    assert _ids(asr, stereo, max_new_tokens=10) == ref
    assert _ids(asr, x.numpy(), max_new_tokens=10) == ref
    assert _ids(asr, x.double(), max_new_tokens=10) == ref


def test_fp32_class_ids_equal_oracle():
    m = ASRModel(SPEC, device="cuda:0", mode="fp32", max_batch=2, max_prompt=320, max_new_tokens=32)
    sd = synthetic_state_dict(ModelDims(enc_layers=2, dec_layers=2), seed=0)
    for x in _segments(2):
        w = x[0].numpy()
        mel, _ = mo.log_mel(mo.prestep(w))
        n_audio = num_audio_tokens(w.shape[0])
        ref, _, _ = ora.generate_greedy(sd, ora.OracleConfig(enc_layers=2, dec_layers=2), torch.from_numpy(mel), n_audio,
                                        synthetic_prompt_ids(n_audio), 16)
        assert _ids(m, x, max_new_tokens=16) == ref
    m.close()


def test_transcribe_batch_equals_single_calls(asr, asr_fp32):
    segs6 = _segments(6)
    assert asr_fp32.transcribe_batch(segs6, max_new_tokens=12) == [asr_fp32.transcribe(s, max_new_tokens=12) for s in segs6]   # exact in fp32
    segs = _segments(6)
    together = asr.transcribe_batch(segs, max_new_tokens=12)
    alone = [asr.transcribe(s, max_new_tokens=12) for s in segs]
    assert [t.split()[:3] for t in together] == [a.split()[:3] for a in alone] and len(together) == 6
    assert [i[:3] for i in asr.transcribe_ids(segs[:3], max_new_tokens=5)] == [_ids(asr, s, max_new_tokens=5)[:3] for s in segs[:3]]


@pytest.fixture(scope="module")
def asr_fp32():
    m = ASRModel(SPEC, device="cuda", mode="fp32", max_batch=8, max_prompt=320, max_new_tokens=64)
    yield m
    m.close()


@pytest.mark.parametrize("n_threads", [4, 16])
@pytest.mark.parametrize("mode", ["fp32", "native"])
def test_concurrent_transcribe_equals_serial(asr, asr_fp32, mode, n_threads):
    """The reference's call pattern: executor threads + the event loop all inside transcribe() on one instance.  fp32 (the
    parity arithmetic, batch-invariant by construction): ids identical to the serial run.  bf16: a segment may be decoded
    by a different batch class / attention split than when it is alone, so ids are identical up to the first step whose
    top-2 margin is inside bf16 noise — the leading tokens and the count must agree."""
    model = asr_fp32 if mode == "fp32" else asr
    segs = _segments(n_threads, base=40)
    serial = [_ids(model, s, max_new_tokens=14) for s in segs]
    before = model.batcher_stats()
    got = [None] * n_threads
    errs = []

    def work(i):
        try:
            for _ in range(3):
                got[i] = _ids(model, segs[i], max_new_tokens=14)
        except Exception as e:           # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(n_threads)]
    [t.start() for t in th]; [t.join() for t in th]
    assert not errs, errs
    if mode == "fp32":
        assert got == serial
    else:
        assert [len(g) for g in got] == [len(s) for s in serial]
        assert [g[:3] for g in got] == [s[:3] for s in serial]
        assert sum(int(g == s) for g, s in zip(got, serial)) >= n_threads // 2
    after = model.batcher_stats()
    assert after["requests"] - before["requests"] == 3 * n_threads
    assert after["batches"] - before["batches"] < 3 * n_threads          # calls were coalesced
    assert after["max_batch_seen"] >= 2


def test_mixed_token_budgets_in_one_batch(asr_fp32):
    """Interim (15 tokens) and committed (min(50+5*dur,200)) requests arriving together keep their own budgets."""
    asr = asr_fp32
    segs = _segments(4, base=80)
    want = [_ids(asr, s, max_new_tokens=g) for s, g in zip(segs, (15, 15, 60, 40))]
    got = [None] * 4

    def work(i, g):
        got[i] = _ids(asr, segs[i], max_new_tokens=g)

    th = [threading.Thread(target=work, args=(i, g)) for i, g in enumerate((15, 15, 60, 40))]
    [t.start() for t in th]; [t.join() for t in th]
    assert got == want and [len(g) for g in got] == [15, 15, 60, 40]


def test_pcm16_entry_point_equals_float_path(asr):
    x = mo.synth_audio("speech", 20480, 5)
    s16 = np.clip(np.rint(x * 32767.0), -32768, 32767).astype(np.int16)
    via_float = asr.transcribe(torch.from_numpy(s16.copy()).float()[None] / 32768.0, max_new_tokens=15)
    assert asr.transcribe_pcm16(s16.tobytes(), max_new_tokens=15) == via_float
    assert asr.transcribe_pcm16(s16, max_new_tokens=15) == via_float


def test_short_window_interim_call(asr_fp32):
    """Opt-in streaming encoder through the drop-in class: explicit per call, and by token budget (SONIC_SHORT_WINDOW_MAX_NEW,
    which serves the reference's unchanged interim call transcribe(..., max_new_tokens=15))."""
    x = (mo.synth_audio("speech", 20480, 21) * 32767).astype(np.int16)
    full = asr_fp32.transcribe_pcm16(x, max_new_tokens=15)
    short = asr_fp32.transcribe_pcm16(x, max_new_tokens=15, short_window=True)
    assert isinstance(short, str) and short
    old = asr_fp32._short_window_max_new
    try:
        asr_fp32._short_window_max_new = 16
        assert asr_fp32.transcribe_pcm16(x, max_new_tokens=15) == short           # by budget
        assert asr_fp32.transcribe_pcm16(x, max_new_tokens=15, short_window=False) == full
        assert asr_fp32.transcribe_pcm16(x, max_new_tokens=32) == asr_fp32.transcribe_pcm16(x, max_new_tokens=32, short_window=False)   # larger budgets keep the full window
    finally:
        asr_fp32._short_window_max_new = old


def test_hotwords_and_argument_errors(asr):
    x = _segments(1)[0]
    a = asr.transcribe(x, max_new_tokens=8, hotwords=["Kubernetes", "B200"])
    assert isinstance(a, str)
    with pytest.raises(ValueError, match="30 s"):
        asr.transcribe(torch.zeros(1, 480001))
    with pytest.raises(ValueError, match="too short"):
        asr.transcribe(torch.zeros(1, 400))
    with pytest.raises(ValueError, match="max_new_tokens"):
        asr.transcribe(x, max_new_tokens=1000)
    with pytest.raises(ValueError):
        ASRModel(SPEC, mode="fp16")
    with pytest.raises(RuntimeError, match="CPU"):
        ASRModel(SPEC, device="cpu")


def test_managers_end_to_end(monkeypatch):
    """models_manager.asr_model_init / asr_model_get + TranscriptionManager on int16 bytes, as connection_manager.py drives
    them (interim: last 20 chunks, 15 tokens; committed: whole segment, min(50 + 5*dur, 200) tokens)."""
    import sonicscribe_b200.models_manager as mm
    import sonicscribe_b200.transcription_manager as tm
    from sonicscribe_b200.config import AppConfig
    monkeypatch.setattr(AppConfig, "CHECKPOINT_PATH", SPEC)
    monkeypatch.setattr(AppConfig, "DEVICE", "cuda")
    monkeypatch.setattr(AppConfig, "SONIC_MAX_BATCH", 4)
    mm.asr_model_reset()
    mm.asr_model_init(max_prompt=320, max_new_tokens=200)
    try:
        model = mm.asr_model_get()
        mm.asr_model_init()                                       # idempotent
        assert mm.asr_model_get() is model
        x = mo.synth_audio("speech", 20 * 1024, 3)
        pcm = np.clip(np.rint(x * 32767.0), -32768, 32767).astype(np.int16).tobytes()
        mgr = tm.TranscriptionManager()
        interim = asyncio.run(mgr.transcribe_temporary(pcm))
        assert interim and len(interim.split()) == 15
        committed = asyncio.run(mgr.transcribe_committed(pcm, 1.28))
        assert len(committed.split()) == 56 and committed.split()[:15] == interim.split()
        assert asyncio.run(mgr.transcribe_temporary(b"\0" * 100)) == ""
    finally:
        mm.asr_model_reset()


def test_release_model_attribute_frees_the_replica():
    """main.py:84-88: ``if hasattr(asr_model, 'model'): del asr_model.model`` at shutdown."""
    m = ASRModel(SPEC, device="cuda", mode="native", max_batch=2, max_prompt=320, max_new_tokens=16)
    x = _segments(1)[0]
    assert m.transcribe(x, max_new_tokens=4)
    free0 = torch.cuda.mem_get_info()[0]
    assert hasattr(m, "model")
    del m.model
    import gc
    gc.collect()
    assert torch.cuda.mem_get_info()[0] > free0 + (100 << 20)
    with pytest.raises(RuntimeError, match="released"):
        m.transcribe(x, max_new_tokens=4)
    m.close()


def test_checkpoint_directory_with_tokenizer_and_chat_template(tmp_path):
    """ASRModel(<dir>): safetensors load, AutoProcessor, chat-template prompt ids, batch_decode — the real-checkpoint
    plumbing, on a locally written directory (tests/stub_checkpoint.py)."""
    from tests.stub_checkpoint import write_checkpoint
    d = write_checkpoint(str(tmp_path / "ckpt"))
    with pytest.warns(UserWarning, match="ignored 1 checkpoint tensors"):
        m = ASRModel(d, device="cuda", mode="native", max_batch=2, max_new_tokens=32)
    assert m.processor is not None and m.config.enc_layers == 2 and m.config.dec_layers == 2
    x = _segments(1)[0]
    info = m.transcribe(x, max_new_tokens=12, hotwords=["Foo", "bar"], return_debug_info=True)
    ids = info["token_ids"]
    assert info["transcript"] == m.processor.batch_decode([ids], skip_special_tokens=True)[0].strip()
    assert "<" not in info["transcript"]                             # real tokenizer text, not the <id> pseudo-text
    # the prompt the engine saw is the processor's own tokenisation of the chat template
    msgs = [{"role": "user", "content": [{"type": "audio", "audio": x[0].numpy()},
                                         {"type": "text", "text": 'Please transcribe this audio into text. Pay special attention to these important terms: "foo", "bar"'}]}]
    hf = m.processor.apply_chat_template(msgs, tokenize=True, add_generation_prompt=True, return_dict=True, return_tensors="pt")
    assert m._prompts.build(num_audio_tokens(x.shape[-1]), ["Foo", "bar"]) == hf["input_ids"][0].tolist()
    # same weights given as a state dict (bf16-rounded like the file) + the same prompt => same ids
    sd = {k: v.to(torch.bfloat16) for k, v in synthetic_state_dict(ModelDims(enc_layers=2, dec_layers=2), seed=0).items()}
    m2 = ASRModel("unused", device="cuda", mode="native", max_batch=2, max_new_tokens=32, state_dict=sd)
    m2._prompts = m._prompts
    assert _ids(m2, x, max_new_tokens=12, hotwords=["Foo", "bar"]) == ids
    m.close(); m2.close()


def test_replica_pool_over_real_replicas():
    """pool.ReplicaPool with real ASRModel replicas on every visible GPU: results in segment order, equal to one replica."""
    from sonicscribe_b200.pool import ReplicaPool, cut_long_segments
    n_gpus = min(torch.cuda.device_count(), 2)
    pool = ReplicaPool(lambda i: ASRModel(SPEC, device=f"cuda:{i}", mode="native", max_batch=4, max_prompt=320, max_new_tokens=16), n_gpus, batch=4)
    try:
        audio = mo.synth_audio("speech", 16000 * 50, 123)
        cuts = cut_long_segments(0, audio.shape[0], 16000, 4.0)
        assert len(cuts) == 13
        segs = [torch.from_numpy(audio[s:e])[None] for s, e in cuts]
        got = pool.transcribe_segments(segs, max_new_tokens=8)
        want = [pool.replicas[0].transcribe(s, max_new_tokens=8) for s in segs]
        assert got == want
    finally:
        pool.close()
