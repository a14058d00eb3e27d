"""Per-phase timing of the persistent decode step (globaltimer stamps after every grid barrier)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sonicscribe_b200.engine import Engine, num_audio_tokens
from sonicscribe_b200.prompt import synthetic_prompt_ids
from sonicscribe_b200.synth import synth_audio
from sonicscribe_b200.weights import ModelDims, synthetic_state_dict
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dims = ModelDims(enc_layers=1, dec_layers=L)
sd = synthetic_state_dict(dims, seed=0)
names = ["qkv_gemm", "attention", "o_gemm+resid", "norm", "gateup+swiglu", "down_gemm", "resid_norm2"]
for B in (1, 16, 64):
    eng = Engine(1, L, mode=os.environ.get("MODE", "bf16"), device=0, max_batch=B, max_prompt=320, max_new=64, debug=True)
    eng.load_state_dict(sd)
    segs = [synth_audio("speech", 320000, seed=i) for i in range(B)]
    prompts = [synthetic_prompt_ids(num_audio_tokens(320000)) for _ in range(B)]
    eng.transcribe_ids(segs, prompts, 24)
    ts = eng.debug_read("persist_ts", 4096)
    d = np.diff(ts)
    per = d[1:1 + 7 * L].reshape(L, 7)
    print(f"B={B}: step {ts[-1]:.1f} us; phase0 {d[0]:.1f}; lm_head {d[1 + 7 * L]:.1f}; pick {d[2 + 7 * L] + d[3 + 7 * L]:.1f}")
    print("   per-layer phase us (median over layers):", {n: round(float(np.median(per[:, i])), 1) for i, n in enumerate(names)}, "layer total", round(float(np.median(per.sum(1))), 1))
    eng.close()
