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
sd = synthetic_state_dict(ModelDims(enc_layers=2, dec_layers=2), seed=0)
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
