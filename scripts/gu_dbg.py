"""Fine-grained stamps of the gate/up phase (layer 1) of the tcgen05 decode class on one CTA (SONIC_PERSIST_DBG_CTA)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sonicscribe_b200.engine import Engine, num_audio_tokens
from sonicscribe_b200.prompt import synthetic_prompt_ids
from sonicscribe_b200.synth import synth_audio
from sonicscribe_b200.weights import ModelDims, synthetic_state_dict
L = 4
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sd = synthetic_state_dict(ModelDims(enc_layers=1, dec_layers=L), seed=0)
names = {0: "prod:start", 1: "prod:item0 issued", 2: "prod:item1 issued", 3: "mma:item0 first stage", 4: "mma:item1 first stage",
         5: "mma:item0 issued", 6: "mma:item1 issued"}
for i in range(2):
    for j, n in enumerate(["acc full", "stores issued", "fence done", "arrived", "summed(last)", "done"]):
        names[8 + 8 * i + j] = f"epi{i}:{n}"
for cta in sys.argv[2:] or ["0", "100", "147"]:
    os.environ["SONIC_PERSIST_DBG_CTA"] = cta
    eng = Engine(1, L, mode=os.environ.get("MODE", "bf16"), device=0, max_batch=B, max_prompt=320, max_new=64, debug=True)
    eng.load_state_dict(sd)
    segs = [synth_audio("speech", 320000, seed=i) for i in range(B)]
    prompts = [synthetic_prompt_ids(num_audio_tokens(320000)) for _ in range(B)]
    eng.transcribe_ids(segs, prompts, 24)
    d = eng.debug_read("rs_dbg", 80)
    dd = eng.debug_read("persist_dbg", 1024)
    print('      converter warp 2, per k block (raw landed, slot free, converted):', np.round(dd[100:130].reshape(10, 3) - dd[0], 2).tolist())
    print(f"cta {cta}: " + "  ".join(f"{names[i]}={d[i]:.2f}" for i in sorted(names) if d[i] >= 0 or i == 0))
    print('      per-warp end of phase', np.round(d[40:56], 2), ' after prefetch', np.round(d[60:76], 2))
    eng.close()
