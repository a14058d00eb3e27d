"""The shipped library really contains the Blackwell instruction classes DESIGN.md claims for each hot kernel
(cuobjdump -sass of sonicscribe_b200/libsonic_b200.so; no GPU needed).  Mnemonics per /opt/skills/guides/B200_PROFILING.md:
UTCHMMA = tcgen05.mma, UTMALDG = TMA tile load, LDTM = tcgen05.ld (TMEM -> registers), FFMA2 / FADD2 = packed fp32x2."""
import os
import re
import shutil
import subprocess
from collections import defaultdict

import pytest

from sonicscribe_b200 import engine

MNEMONICS = ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "FFMA2", "FADD2", "HMMA", "SHFL", "MUFU.EX2", "MUFU.LG2")


@pytest.fixture(scope="module")
def sass_counts():
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(engine.lib_path()):
        pytest.skip("cuobjdump or the built library is not available")
    txt = subprocess.run([exe, "-sass", engine.lib_path()], capture_output=True, text=True, check=True).stdout
    counts, cur = defaultdict(lambda: defaultdict(int)), None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        for mn in MNEMONICS:
            if " " + mn in line:
                counts[cur][mn] += 1
    return counts


def _kernels(counts, *needles):
    return {k: v for k, v in counts.items() if all(n in k for n in needles)}


def test_gemm_and_attention_run_on_tcgen05_fed_by_tma(sass_counts):
    for needles in (("gemm_tc_persist_kernel",), ("attention_tc_kernel",), ("attention_prefill_tc_kernel",)):
        ks = _kernels(sass_counts, *needles)
        assert ks, needles
        for name, c in ks.items():
            assert c["UTCHMMA"] > 0 and c["UTMALDG"] > 0 and c["LDTM"] > 0, (name, dict(c))
    enc_attn = next(iter(_kernels(sass_counts, "attention_tc_kernel").values()))
    assert enc_attn["FFMA2"] > 0 and enc_attn["FADD2"] > 0            # packed fp32x2 softmax arithmetic


def test_decode_step_classes(sass_counts):
    ks = _kernels(sass_counts, "decode_persist_kernel")
    assert len(ks) == 10                                               # 8 register-streaming classes + 2 tcgen05 classes
    tc = {k: v for k, v in ks.items() if k.endswith("Lb1EEEvNS_17DecodePersistArgsE")}
    assert len(tc) == 2                                                # bf16 and int8 weights
    for name, c in tc.items():
        assert c["UTCHMMA"] > 0 and c["UTMALDG"] > 0 and c["LDTM"] > 0, (name, dict(c))
    for name, c in ks.items():
        if name not in tc:
            assert c["UTCHMMA"] == 0 and c["HMMA"] > 0 and c["UTMALDG"] > 0, (name, dict(c))     # mma.sync GEMMs, TMA K/V tiles


def test_log_mel_uses_packed_fp32_and_shuffles(sass_counts):
    ks = _kernels(sass_counts, "mel_frames_kernel")
    assert len(ks) == 2                                                # float and bf16 time-major copies
    for name, c in ks.items():
        assert c["FFMA2"] >= 100 and c["FADD2"] >= 40 and c["SHFL"] >= 50 and c["MUFU.LG2"] > 0, (name, dict(c))
        assert c["UTCHMMA"] == 0                                       # the front end is not reshaped into a tensor-core GEMM
