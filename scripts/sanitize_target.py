"""Target of scripts/sanitize.sh: a short pass over every kernel family on the tiny (2+2 layer) model — the smoke path
(fp32 + bf16, one segment) and persistent decode steps at 64 / 20 / 3 segments (bf16) and 2 / 20 segments (int8: register-streaming class / tcgen05 class with converter warps)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from sonicscribe_b200.engine import Engine, num_audio_tokens  # noqa: E402
from sonicscribe_b200.prompt import synthetic_prompt_ids  # noqa: E402
from sonicscribe_b200.synth import synth_audio  # noqa: E402
from sonicscribe_b200.weights import ModelDims, synthetic_state_dict  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
steps = int(os.environ.get("SANITIZE_TOKENS", "3"))
if what == "mel":
    # the log-mel kernels alone (no weights needed): ragged float32 segments incl. the 30 s maximum and one shorter than a frame
    # tile, the int16 wire format, features only / time-major copy, both element types of the time-major copy
    from sonicscribe_b200.engine import FLAG_FEATURES_ONLY, FLAG_PCM_S16, FLAG_REFERENCE_PRESTEP  # noqa: E402
    lens = [20480, 479999, 480000, 1600, 48000, 163841]
    segs = [synth_audio("speech" if i % 2 else "noise", n, seed=i) for i, n in enumerate(lens)]
    for mode in ("bf16", "fp32"):
        eng = Engine(1, 1, mode=mode, device=0, max_batch=len(lens), max_prompt=32, max_new=2)
        f1, nf = eng.mel(segs)
        f2, _ = eng.mel(segs, flags=FLAG_REFERENCE_PRESTEP | FLAG_FEATURES_ONLY)
        assert np.array_equal(f1, f2) and np.isfinite(f1).all()
        eng.mel([(s * 32767).astype(np.int16) for s in segs[:3]], flags=FLAG_REFERENCE_PRESTEP | FLAG_PCM_S16, want_features=False)
        print(f"[sanitize-target] mel: {mode} ok, frames {nf.tolist()}, launches {eng.launch_count()}", flush=True)
        eng.close()
    sys.exit(0)
sd = synthetic_state_dict(ModelDims(enc_layers=2, dec_layers=2), seed=0)
if what == "short":
    # opt-in short encoder window (SONIC_FLAG_SHORT_WINDOW): two 1.28 s segments, then a full-window call on the same handle
    from sonicscribe_b200.engine import FLAG_REFERENCE_PRESTEP, FLAG_SHORT_WINDOW  # noqa: E402
    segs = [synth_audio("speech", 20480, seed=i) for i in range(2)]
    prompts = [synthetic_prompt_ids(num_audio_tokens(20480))] * 2
    eng = Engine(2, 2, mode="bf16", device=0, max_batch=2, max_prompt=64, max_new=8)
    eng.load_state_dict(sd)
    a = eng.transcribe_ids(segs, prompts, steps, flags=FLAG_REFERENCE_PRESTEP | FLAG_SHORT_WINDOW)
    b = eng.transcribe_ids(segs, prompts, steps)
    assert all(len(o) == steps for o in a + b)
    print(f"[sanitize-target] short: bf16 ok {a} / full {b}, launches {eng.launch_count()}", flush=True)
    eng.close()
    sys.exit(0)
cases = {"smoke_fp32": ("fp32", 1), "smoke_bf16": ("bf16", 1), "b64": ("bf16", 64), "b20": ("bf16", 20), "b3": ("bf16", 3), "int8_b2": ("int8", 2), "int8_b20": ("int8", 20)}
for name, (mode, B) in cases.items():
    if what not in ("all", name):
        continue
    lens = [32000 + 1600 * (i % 5) for i in range(B)]
    segs = [synth_audio("speech", n, seed=i) for i, n in enumerate(lens)]
    prompts = [synthetic_prompt_ids(num_audio_tokens(n)) for n in lens]
    eng = Engine(2, 2, mode=mode, device=0, max_batch=B, max_prompt=64, max_new=8)
    eng.load_state_dict(sd)
    out = eng.transcribe_ids(segs, prompts, steps)
    assert all(len(o) == steps and all(0 <= t < 59264 for t in o) for o in out)
    print(f"[sanitize-target] {name}: {mode} B={B} ok, launches {eng.launch_count()}", flush=True)
    eng.close()
