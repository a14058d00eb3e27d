"""CPU-side checks: the C-ABI library loads, exports every symbol include/sonic_b200.h declares, and fails loudly
without a GPU (no compute calls here)."""
import os
import re

import pytest

from sonicscribe_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "sonic_b200.h")).read()
    return sorted(set(re.findall(r"SONIC_API[^;]*?\b(sonic_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(engine.lib_path()):
        import __graft_entry__ as g
        g.build()
    lib = engine.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(lib._sonic_protos) == declared
    assert b"sm_100a" in lib.sonic_version()


def test_token_count_matches_c_side():
    lib = engine.load_library()
    for n in (1600, 16000, 20480, 319963, 320000, 479999, 480000, 600000):
        assert lib.sonic_num_audio_tokens(n) == engine.num_audio_tokens(n)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device|CPU"):
        engine.Engine(2, 2, mode="fp32")
    from sonicscribe_b200.asr import ASRModel
    with pytest.raises(RuntimeError):
        ASRModel("synthetic:enc=2,dec=2", device="cpu")
    with pytest.raises(ValueError):
        ASRModel("synthetic", mode="fp16")


def test_flag_values_match_the_header():
    """engine.py mirrors the SONIC_FLAG_* bits of include/sonic_b200.h by value."""
    txt = open(os.path.join(ROOT, "include", "sonic_b200.h")).read()
    header = {m.group(1): int(m.group(2), 16) for m in re.finditer(r"#define\s+SONIC_FLAG_(\w+)\s+0x([0-9a-fA-F]+)", txt)}
    mirror = {"PEAK_NORM": engine.FLAG_PEAK_NORM, "PCM16": engine.FLAG_PCM16, "PCM_S16": engine.FLAG_PCM_S16,
              "PCM_DEVICE": engine.FLAG_PCM_DEVICE, "OUT_DEVICE": engine.FLAG_OUT_DEVICE, "FEATURES_ONLY": engine.FLAG_FEATURES_ONLY,
              "SHORT_WINDOW": engine.FLAG_SHORT_WINDOW}
    assert header == mirror
    assert len(set(mirror.values())) == len(mirror) and all(v & (v - 1) == 0 for v in mirror.values())     # distinct single bits
    assert engine.FLAG_REFERENCE_PRESTEP == engine.FLAG_PEAK_NORM | engine.FLAG_PCM16
