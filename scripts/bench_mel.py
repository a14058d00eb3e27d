"""BASELINE.json configs[4]: log-mel front-end bandwidth sweep, 1-1024 concurrent 20 s segments, device-resident PCM.
Algorithmic bytes per segment: 1 280 000 B of fp32 PCM read + 1 536 000 B of fp32 features written."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
from sonicscribe_b200.engine import Engine, FLAG_PCM_DEVICE, FLAG_OUT_DEVICE, FLAG_REFERENCE_PRESTEP, FLAG_FEATURES_ONLY
from sonicscribe_b200.synth import synth_audio

N = 320000
# MEL_WITH_TM=1: also write the encoder's time-major bf16 copy (what sonic_mel does inside the transcription path)
EXTRA = 0 if os.environ.get('MEL_WITH_TM') else FLAG_FEATURES_ONLY
maxB = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
eng = Engine(1, 1, mode="bf16", device=0, max_batch=maxB, max_prompt=32, max_new=2)
base = np.stack([synth_audio("speech", N, seed=i) for i in range(8)])
pcm = torch.from_numpy(np.tile(base, (maxB // 8 + 1, 1))[:maxB].copy()).cuda()
feat = torch.empty((maxB, 128, 3000), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
peak = 6542.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
rows = []
B = int(os.environ.get('MEL_ONLY_B', '1'))
while B <= maxB:
    offs = (np.arange(B, dtype=np.int64) * N)
    lens = np.full(B, N, dtype=np.int32)
    nfr = np.zeros(B, dtype=np.int32)
    def call():
        rc = eng.lib.sonic_mel(eng.h, C.c_void_p(pcm.data_ptr()), offs.ctypes.data_as(C.POINTER(C.c_int64)), lens.ctypes.data_as(C.POINTER(C.c_int32)),
                               B, FLAG_REFERENCE_PRESTEP | FLAG_PCM_DEVICE | FLAG_OUT_DEVICE | EXTRA, C.c_void_p(feat.data_ptr()), nfr.ctypes.data_as(C.POINTER(C.c_int32)))
        assert rc == 0, eng.lib.sonic_last_error(eng.h)
    for _ in range(3):
        call()
    iters = max(3, min(50, 2048 // B)) if 'MEL_ONLY_B' not in os.environ else 3
    eng.timer_begin()
    for _ in range(iters):
        call()
    ms = eng.timer_end() / iters
    gbs = B * 2816000 / (ms * 1e-3) / 1e9
    rows.append({"segments": B, "ms": ms, "GBps": gbs, "frac_of_measured_hbm_peak": gbs / peak, "audio_s_per_s": B * 20.0 / (ms * 1e-3)})
    print(rows[-1], flush=True)
    B *= 2
print(json.dumps({"metric": "log-mel front end, achieved algorithmic GB/s (pre-step + STFT + mel + log + clamp, fp32 features out)", "peak_GBps": peak, "sweep": rows}))
