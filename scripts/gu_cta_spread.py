"""Per-CTA start / work-done / barrier-passed times of the gate/up phase of layer 1 (SONIC_PERSIST_DBG_CTA=-2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
os.environ["SONIC_PERSIST_DBG_CTA"] = "-2"
from sonicscribe_b200.engine import Engine, num_audio_tokens
from sonicscribe_b200.prompt import synthetic_prompt_ids
from sonicscribe_b200.synth import synth_audio
from sonicscribe_b200.weights import ModelDims, synthetic_state_dict
L = 4
sd = synthetic_state_dict(ModelDims(enc_layers=1, dec_layers=L), seed=0)
for B in [int(b) for b in sys.argv[1:]] or [64, 1]:
    eng = Engine(1, L, mode=os.environ.get("MODE", "bf16"), device=0, max_batch=B, max_prompt=320, max_new=64, debug=True)
    eng.load_state_dict(sd)
    segs = [synth_audio("speech", 320000, seed=i) for i in range(B)]
    prompts = [synthetic_prompt_ids(num_audio_tokens(320000)) for _ in range(B)]
    eng.transcribe_ids(segs, prompts, 24)
    d = eng.debug_read("persist_dbg", 1024)
    n = 148
    st, dn, ps = d[:n], d[160:160 + n], d[320:320 + n]
    q = lambda x: "min %.2f p50 %.2f p90 %.2f max %.2f" % (x.min(), np.median(x), np.percentile(x, 90), x.max())
    print(f"B={B}: start [{q(st)}]  work done [{q(dn)}]  barrier passed [{q(ps)}]")
    print(f"   barrier: tid0 enters [{q(d[480:480 + n])}]  arrival performed [{q(d[640:640 + n])}]  poll succeeded [{q(d[800:800 + n])}]")
    print("   CTA %d: per-warp barrier entry " % (int(os.environ.get("SONIC_PERSIST_DBGFLAGS", "0")) >> 8), np.round(d[960:976], 2), " after proxy fence ", np.round(d[980:996], 2), " after bar.sync ", np.round(d[1000:1016], 2))
    en = d[480:480 + n]
    oe = np.argsort(en)
    print("   barrier entry, slowest CTAs:", [(int(i), round(float(en[i]), 2)) for i in oe[-12:]], " histogram (us from 7 to 16):", np.histogram(en, bins=np.arange(7, 17))[0])
    order = np.argsort(dn)
    print("   slowest CTAs (id: done):", [(int(i), round(float(dn[i]), 2)) for i in order[-8:]], " fastest:", [(int(i), round(float(dn[i]), 2)) for i in order[:4]])
    eng.close()
