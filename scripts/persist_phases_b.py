"""Per-phase timing of the persistent decode step at chosen batch sizes: python persist_phases_b.py L B [B ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sonicscribe_b200.engine import Engine, num_audio_tokens
from sonicscribe_b200.prompt import synthetic_prompt_ids
from sonicscribe_b200.synth import synth_audio
from sonicscribe_b200.weights import ModelDims, synthetic_state_dict
L = int(sys.argv[1])
sd = synthetic_state_dict(ModelDims(enc_layers=1, dec_layers=L), seed=0)
names = ["qkv", "attn", "o", "norm", "gateup", "down", "norm2"]
for B in [int(b) for b in sys.argv[2:]]:
    eng = Engine(1, L, mode=os.environ.get("MODE", "bf16"), device=0, max_batch=B, max_prompt=320, max_new=64, debug=True)
    eng.load_state_dict(sd)
    segs = [synth_audio("speech", 320000, seed=i) for i in range(B)]
    prompts = [synthetic_prompt_ids(num_audio_tokens(320000)) for _ in range(B)]
    eng.transcribe_ids(segs, prompts, 24)
    ts = eng.debug_read("persist_ts", 4096)
    d = np.diff(ts)
    per = d[1:1 + 7 * L].reshape(L, 7)
    print(f"B={B}: step {ts[-1]:.1f} phase0 {d[0]:.1f} lm_head {d[1 + 7 * L]:.1f} pick {d[2 + 7 * L] + d[3 + 7 * L]:.1f} | " +
          " ".join(f"{n} {np.median(per[:, i]):.1f}" for i, n in enumerate(names)) + f" | layer {np.median(per.sum(1)):.1f}", flush=True)
    eng.close()
