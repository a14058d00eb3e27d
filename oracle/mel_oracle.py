"""CPU oracle for the audio pre-step and the 128-bin log-mel front end.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``sonicscribe_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may use it, and only as
the checker / baseline.

The reference (``/root/reference/backend/asr.py``) contains no arithmetic of its
own for this stage: it calls the third-party ``transformers`` package
(unpinned in ``backend/requirements.txt:12``; this oracle is pinned to the
container's transformers 5.5.0).  This file restates, in plain numpy float64:

* the pre-step of ``backend/asr.py:248-276`` (first channel, peak-normalise,
  PCM_16 WAV round trip through ``soundfile``),
* ``WhisperFeatureExtractor._torch_extract_fbank_features``
  (``transformers/models/whisper/feature_extraction_whisper.py:135-164``),
  padding/mask handling (``:296-337``) and
* the slaney mel filter bank (``transformers/audio_utils.py:263-332,356-375,
  453-544``).

Parity status: PINNED against the importable HF implementation by
``tests/golden/gen_golden.py`` (fixtures ``tests/golden/mel_*.npz``) and against
the survey's bootstrap checksums.  The PCM_16 round-trip
constant (write scale 32767? 32768?) follows libsndfile's documented float->short
conversion and is NOT pinned by a run of ``soundfile`` (package absent here) —
see ``pcm16_roundtrip``.
"""
from __future__ import annotations

import numpy as np

SAMPLE_RATE = 16000
N_FFT = 400
HOP = 160
N_MELS = 128
N_SAMPLES = 480000          # 30 s window  (feature_extraction_whisper.py:45-49, chunk_length=30)
N_FRAMES = N_SAMPLES // HOP  # 3000
N_BINS = N_FFT // 2 + 1      # 201


# ----------------------------------------------------------------------------------------------
# pre-step  (backend/asr.py:248-276)
# ----------------------------------------------------------------------------------------------
def peak_normalise(x: np.ndarray) -> np.ndarray:
    """``wav / max|wav|`` iff ``max|wav| > 1e-6``  (backend/asr.py:265-267), float32 arithmetic."""
    x = np.asarray(x, dtype=np.float32)
    m = np.float32(np.max(np.abs(x))) if x.size else np.float32(0)
    if m > np.float32(1e-6):
        x = (x / m).astype(np.float32)
    return x


def pcm16_roundtrip(x: np.ndarray) -> np.ndarray:
    """``soundfile.write(path, x, sr)`` (default subtype PCM_16, backend/asr.py:276) followed by the
    processor's float32 re-load (transformers/audio_utils.py:60-88).

    libsndfile float->short without clipping enabled: ``lrintf(x * 0x7FFF)``; short->float on read:
    ``s / 0x8000``.  lrintf = round-half-even = ``np.rint``.  After peak normalisation |x|<=1 so the
    product never exceeds 32767 and no clipping is involved.
    """
    x = np.asarray(x, dtype=np.float32)
    q = np.rint(x.astype(np.float32) * np.float32(32767.0))
    q = np.clip(q, -32768, 32767)
    return (q / np.float32(32768.0)).astype(np.float32)


def prestep(audio: np.ndarray, peak_norm: bool = True, pcm16: bool = True) -> np.ndarray:
    """Full pre-step of ``ASRModel._prepare_audio_tempfile`` for 16 kHz input: [C,N] or [N] -> [N]."""
    a = np.asarray(audio, dtype=np.float32)
    if a.ndim == 2:
        a = a[0]
    if peak_norm:
        a = peak_normalise(a)
    if pcm16:
        a = pcm16_roundtrip(a)
    return a


# ----------------------------------------------------------------------------------------------
# mel filter bank  (transformers/audio_utils.py:453-544, slaney scale + slaney norm)
# ----------------------------------------------------------------------------------------------
def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    mel = 3.0 * f / 200.0
    logstep = 27.0 / np.log(6.4)
    hi = f >= 1000.0
    mel = np.where(hi, 15.0 + np.log(np.maximum(f, 1e-30) / 1000.0) * logstep, mel)
    return mel


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f = 200.0 * m / 3.0
    logstep = np.log(6.4) / 27.0
    hi = m >= 15.0
    f = np.where(hi, 1000.0 * np.exp(logstep * (m - 15.0)), f)
    return f


def mel_filter_bank() -> np.ndarray:
    """[201, 128] float64 filter bank; cast to float32 by the caller (feature_extraction_whisper.py:152)."""
    mel_pts = np.linspace(_hz_to_mel_slaney(0.0), _hz_to_mel_slaney(8000.0), N_MELS + 2)
    hz_pts = _mel_to_hz_slaney(mel_pts)
    fft_freqs = np.linspace(0, SAMPLE_RATE // 2, N_BINS)
    diff = np.diff(hz_pts)
    slopes = hz_pts[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / diff[:-1]
    up = slopes[:, 2:] / diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (hz_pts[2:N_MELS + 2] - hz_pts[:N_MELS])
    return fb * enorm[None, :]


def sparse_mel_taps(max_taps: int = 12):
    """(start[128] int32, count[128] int32, weights[128, max_taps] float32): the non-zero band of each filter."""
    fb = mel_filter_bank().astype(np.float32)
    start = np.zeros(N_MELS, np.int32)
    count = np.zeros(N_MELS, np.int32)
    w = np.zeros((N_MELS, max_taps), np.float32)
    for m in range(N_MELS):
        nz = np.nonzero(fb[:, m])[0]
        lo, hi = int(nz[0]), int(nz[-1])
        assert hi - lo + 1 <= max_taps
        start[m], count[m] = lo, hi - lo + 1
        w[m, : hi - lo + 1] = fb[lo : hi + 1, m]
    return start, count, w


# ----------------------------------------------------------------------------------------------
# log-mel  (feature_extraction_whisper.py:135-164, 296-337)
# ----------------------------------------------------------------------------------------------
def n_valid_frames(n: int) -> int:
    """``attention_mask[:, ::160]`` summed: ceil(min(n,480000)/160)  (feature_extraction_whisper.py:328-337)."""
    n = min(int(n), N_SAMPLES)
    return -(-n // HOP)


def n_audio_tokens(n: int) -> int:
    """processing_glmasr.py:97-103 / modeling_glmasr.py:417-421 on the frame mask length."""
    f = n_valid_frames(n)
    c = (f - 1) // 2 + 1          # conv2 (k3,s2,p1); conv1 keeps the length
    return (c - 4) // 4 + 1


def log_mel(x: np.ndarray, dtype=np.float64):
    """x: [N] float32 waveform (after the pre-step).  Returns (features [128,3000] float32, mask [3000] int32).

    Steps: zero-pad/truncate to 480000, reflect-pad 200, 3001 frames of 400 with
    periodic Hann, rFFT, |.|^2, drop the last frame, mel projection, log10(clamp 1e-10), max(., gmax-8), (.+4)/4.
    """
    x = np.asarray(x, dtype=np.float32)
    n = min(x.shape[0], N_SAMPLES)
    y = np.zeros(N_SAMPLES, dtype=dtype)
    y[:n] = x[:n]
    ypad = np.pad(y, (N_FFT // 2, N_FFT // 2), mode="reflect")
    win = (0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(N_FFT) / N_FFT)).astype(np.float32).astype(dtype)
    idx = np.arange(N_FFT)[None, :] + HOP * np.arange(N_FRAMES)[:, None]
    frames = ypad[idx] * win[None, :]
    spec = np.fft.rfft(frames.astype(np.float64), axis=1)
    power = (spec.real ** 2 + spec.imag ** 2).astype(dtype)          # [3000, 201]
    fb = mel_filter_bank().astype(np.float32).astype(dtype)           # [201, 128]
    mel = power @ fb                                                  # [3000, 128]
    logm = np.log10(np.maximum(mel, 1e-10))
    g = logm.max()
    logm = np.maximum(logm, g - 8.0)
    out = ((logm + 4.0) / 4.0).T.astype(np.float32)                   # [128, 3000]
    mask = (HOP * np.arange(N_FRAMES) < n).astype(np.int32)
    return np.ascontiguousarray(out), mask


def log_mel_from_audio(audio: np.ndarray, peak_norm: bool = True, pcm16: bool = True):
    return log_mel(prestep(audio, peak_norm, pcm16))


# ----------------------------------------------------------------------------------------------
# deterministic synthetic audio families: single definition in sonicscribe_b200/synth.py (input
# generation only, no arithmetic of the path), re-exported here for the tests.
# ----------------------------------------------------------------------------------------------
from sonicscribe_b200.synth import synth_audio  # noqa: E402,F401
