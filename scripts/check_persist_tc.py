"""tcgen05 GEMM phases of the persistent decode kernel (batch class 33..64) against its mma.sync phases: same token ids."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sonicscribe_b200.engine import Engine, num_audio_tokens
from sonicscribe_b200.prompt import synthetic_prompt_ids
from sonicscribe_b200.synth import synth_audio
from sonicscribe_b200.weights import ModelDims, synthetic_state_dict
L = int(sys.argv[1]) if len(sys.argv) > 1 else 4
G = int(sys.argv[2]) if len(sys.argv) > 2 else 24
dims = ModelDims(enc_layers=1, dec_layers=L)
sd = synthetic_state_dict(dims, seed=0)
for B in (40, 64):
    segs = [synth_audio("speech", 320000 - 4000 * (i % 7), seed=i) for i in range(B)]
    prompts = [synthetic_prompt_ids(num_audio_tokens(len(s))) for s in segs]
    out = {}
    for tc in ("0", "1"):
        os.environ["SONIC_PERSIST_TC"] = tc
        eng = Engine(1, L, mode="bf16", device=0, max_batch=B, max_prompt=320, max_new=64, debug=True)
        eng.load_state_dict(sd)
        ids, mar = eng.transcribe_ids(segs, prompts, G, want_margins=True)
        ts = eng.debug_read("persist_ts", 4096)
        out[tc] = (ids, ts[-1], mar)
        eng.close()
    same = sum(int(np.array_equal(a, b)) for a, b in zip(out["0"][0], out["1"][0]))
    first_div = [next((i for i, (x, y) in enumerate(zip(a, b)) if x != y), -1) for a, b in zip(out["0"][0], out["1"][0])]
    worst = 0.0; mdiff = 0.0
    for s_, d in enumerate(first_div):
        n = d if d >= 0 else G
        if d >= 0: worst = max(worst, min(out["0"][2][s_][d], out["1"][2][s_][d]))
        if n: mdiff = max(mdiff, float(np.abs(np.array(out["0"][2][s_][:n]) - np.array(out["1"][2][s_][:n])).max()))
    print(f"   largest top-2 margin at a first divergence {worst:.4f}; max |margin difference| before divergence {mdiff:.4f}")
    print(f"B={B}: identical id sequences {same}/{B}; step us mma {out['0'][1]:.1f} tc {out['1'][1]:.1f}; first divergences {[d for d in first_div if d >= 0][:8]}", flush=True)
