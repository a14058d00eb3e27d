"""Stamps of the split-KV attention phase (layer 1) on one CTA: python attn_dbg.py B cta [cta ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sonicscribe_b200.engine import Engine, num_audio_tokens
from sonicscribe_b200.prompt import synthetic_prompt_ids
from sonicscribe_b200.synth import synth_audio
from sonicscribe_b200.weights import ModelDims, synthetic_state_dict
L = 4
B = int(sys.argv[1])
sd = synthetic_state_dict(ModelDims(enc_layers=1, dec_layers=L), seed=0)
for cta in sys.argv[2:]:
    os.environ["SONIC_PERSIST_DBG_CTA"] = cta
    eng = Engine(1, L, mode=os.environ.get("MODE", "bf16"), device=0, max_batch=B, max_prompt=320, max_new=64, debug=True)
    eng.load_state_dict(sd)
    segs = [synth_audio("speech", 320000, seed=i) for i in range(B)]
    prompts = [synthetic_prompt_ids(num_audio_tokens(320000)) for _ in range(B)]
    eng.transcribe_ids(segs, prompts, 24)
    d = eng.debug_read("persist_dbg", 1024)
    ts = eng.debug_read("persist_ts", 4096)
    print(f"cta {cta}: attention stamps (start, kv loads issued, preamble, kv in smem, scores, softmax, pv+ws / arrive, last?, merge..):", np.round(d[900:916], 2), " phase total", round(float(np.diff(ts)[1 + 7 + 1]), 2))
    eng.close()
