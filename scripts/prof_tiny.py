"""Short workload for ncu: full-width 2+2 layer model, B segments of 20 s, a few greedy steps (same kernel shapes as the
full model, 1/16 of the launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sonicscribe_b200.engine import Engine, num_audio_tokens
from sonicscribe_b200.prompt import synthetic_prompt_ids
from sonicscribe_b200.synth import synth_audio
from sonicscribe_b200.weights import ModelDims, synthetic_state_dict

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
G = int(sys.argv[2]) if len(sys.argv) > 2 else 4
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dims = ModelDims(enc_layers=2, dec_layers=2)
eng = Engine(2, 2, mode="bf16", device=0, max_batch=B, max_prompt=320, max_new=128)
eng.load_state_dict(synthetic_state_dict(dims, seed=0))
segs = [synth_audio("speech", 320000, seed=i) for i in range(B)]
prompts = [synthetic_prompt_ids(num_audio_tokens(320000)) for _ in range(B)]
for _ in range(reps):
    out = eng.transcribe_ids(segs, prompts, G)
print("done", out[0])
