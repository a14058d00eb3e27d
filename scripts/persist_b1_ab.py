"""A/B of the per-token decode kernel at small batches: full 28-layer step time (persist_ts stamps), several repetitions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sonicscribe_b200.engine import Engine, num_audio_tokens
from sonicscribe_b200.prompt import synthetic_prompt_ids
from sonicscribe_b200.synth import synth_audio
from sonicscribe_b200.weights import ModelDims, synthetic_state_dict
L = 8
dims = ModelDims(enc_layers=1, dec_layers=L)
sd = synthetic_state_dict(dims, seed=0)
names = ["qkv", "attn", "o", "norm", "gateup", "down", "norm2"]
for B in (1, 16):
    eng = Engine(1, L, mode="bf16", device=0, max_batch=B, max_prompt=320, max_new=64, debug=True)
    eng.load_state_dict(sd)
    segs = [synth_audio("speech", 320000, seed=i) for i in range(B)]
    prompts = [synthetic_prompt_ids(num_audio_tokens(320000)) for _ in range(B)]
    steps = []
    for rep in range(4):
        eng.transcribe_ids(segs, prompts, 24)
        ts = eng.debug_read("persist_ts", 4096)
        steps.append(float(ts[-1]))
    d = np.diff(ts)
    per = d[1:1 + 7 * L].reshape(L, 7)
    print(f"B={B}: step us {np.round(steps, 1).tolist()}; per-layer {dict(zip(names, np.round(np.median(per, 0), 1).tolist()))} lm_head {d[1 + 7 * L]:.1f}", flush=True)
    eng.close()
