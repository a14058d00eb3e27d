"""Phase timestamps of the row-sliced decode step (decode_rs.cu) on the full-size model: us per phase, averaged over layers.
usage: python scripts/rs_phases.py [mode] [B ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sonicscribe_b200.engine import Engine, num_audio_tokens
from sonicscribe_b200.prompt import synthetic_prompt_ids
from sonicscribe_b200.synth import synth_audio
from sonicscribe_b200.weights import ModelDims, iter_synthetic_tensors

mode = sys.argv[1] if len(sys.argv) > 1 else "bf16"
Bs = [int(v) for v in sys.argv[2:]] or [1, 16]
L = int(os.environ.get("RS_LAYERS", "28"))
eng = Engine(2, L, mode=mode, device=0, max_batch=max(Bs), max_prompt=320, max_new=64, debug=True)
eng.load_state_dict(iter_synthetic_tensors(ModelDims(enc_layers=2, dec_layers=L), seed=0))
for B in Bs:
    segs = [synth_audio("speech", 320000, seed=i) for i in range(B)]
    prompts = [synthetic_prompt_ids(num_audio_tokens(320000)) for _ in range(B)]
    eng.transcribe_ids(segs, prompts, 48)
    ts = eng.debug_read("rs_ts", 5 * L + 3)
    d = np.diff(ts)
    per = d[: 5 * L].reshape(L, 5)
    names = ["qkv+rope", "attention", "o+resid", "gate/up", "down+resid"]
    print(f"{mode} B={B}: step {ts[-1]:.1f} us; layer {per.sum(1).mean():.2f} us; " +
          "; ".join(f"{n} {v:.2f}" for n, v in zip(names, per.mean(0))) + f"; lm_head {d[5 * L]:.1f}; pick {d[5 * L + 1]:.1f}", flush=True)
    print("   stage ms", eng.stage_times(), flush=True)
    dbg = eng.debug_read("rs_dbg", 80).reshape(5, 16)
    pts = ["start", "built", "mma:begin", "mma:first-stage", "mma:issued", "epi:acc-full", "epi:done", "bar:enter", "bar:arrived", "bar:released",
           "prod:issued", "act:issued"]
    for ph, nm in enumerate(["qkv", "attn", "o", "gate/up", "down"]):
        base = dbg[ph][0]
        print(f"   dbg {nm:8s} " + " ".join(f"{p}={dbg[ph][i] - base:7.2f}" for i, p in enumerate(pts) if dbg[ph][i] >= 0), flush=True)
eng.close()
