import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sonicscribe_b200.engine import Engine, num_audio_tokens
from sonicscribe_b200.prompt import synthetic_prompt_ids
from sonicscribe_b200.synth import synth_audio
from sonicscribe_b200.weights import ModelDims, synthetic_state_dict
B = int(sys.argv[1]); L = int(sys.argv[2]); G = int(sys.argv[3])
dims = ModelDims(enc_layers=1, dec_layers=L)
eng = Engine(1, L, mode="bf16", device=0, max_batch=B, max_prompt=int(os.environ.get("MP", "320")), max_new=int(os.environ.get("MN", "64")), debug=os.environ.get("DBG", "0") == "1")
eng.load_state_dict(synthetic_state_dict(dims, seed=0))
segs = [synth_audio("speech", 320000, seed=i) for i in range(B)]
prompts = [synthetic_prompt_ids(num_audio_tokens(320000)) for _ in range(B)]
print(eng.transcribe_ids(segs, prompts, G)[0])
