"""Generate the golden fixtures under tests/golden/ from the REAL third-party implementation the reference
calls (transformers 5.5.0: WhisperFeatureExtractor, GlmAsrForConditionalGeneration.generate) — run in the build
container only (CPU, ~3 min).  The GPU box never runs this; it reads the committed .npz files.

    python tests/golden/gen_golden.py            # all fixtures
    python tests/golden/gen_golden.py mel tiny   # a subset

Inputs are fully determined by (kind, n_samples, seed) via oracle.mel_oracle.synth_audio and by
sonicscribe_b200.weights.synthetic_state_dict(dims, seed), so only HF *outputs* are stored.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import mel_oracle as mo  # noqa: E402
from sonicscribe_b200.weights import ModelDims, synthetic_state_dict  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

MEL_CASES = [  # (kind, n_samples, seed, prestep)
    ("noise", 320000, 0, False),
    ("noise", 320000, 0, True),
    ("speech", 320000, 1, True),
    ("speech", 319963, 2, True),
    ("speech", 20480, 3, True),
    ("noise", 1600, 4, True),
    ("noise", 479999, 5, True),
    ("noise", 480000, 6, True),
    ("speech", 500000, 7, True),     # longer than the 30 s window: truncated
    ("zeros", 32000, 0, True),
    ("impulse", 48000, 0, True),
    ("square", 16000, 0, True),
]
FRAME_STRIDE = 5


def synthetic_prompt_ids(n_audio: int):
    """No tokenizer/chat template in the container: fixed synthetic prefix(8)/suffix(12)."""
    return list(range(100, 108)) + [59260] * n_audio + list(range(200, 212))


def hf_features(x):
    from transformers import WhisperFeatureExtractor

    fe = WhisperFeatureExtractor(feature_size=128)
    f = fe(x, sampling_rate=16000, return_attention_mask=True, padding="max_length", return_tensors="pt")
    return f["input_features"][0], f["attention_mask"][0]


def gen_mel():
    out = {}
    for i, (kind, n, seed, pre) in enumerate(MEL_CASES):
        x = mo.synth_audio(kind, n, seed)
        xp = mo.prestep(x) if pre else x
        feat, mask = hf_features(xp)
        feat = feat.numpy()
        out[f"c{i}_sub"] = feat[:, ::FRAME_STRIDE].copy()
        out[f"c{i}_stats"] = np.array([feat.astype(np.float64).sum(), feat.max(), feat.min(), int(mask.sum())])
    np.savez_compressed(os.path.join(OUT, "mel_cases.npz"), **out)
    print("wrote mel_cases.npz")


def build_hf(dims: ModelDims, sd):
    from transformers import GlmAsrConfig, GlmAsrForConditionalGeneration

    cfg = GlmAsrConfig(audio_config={"num_hidden_layers": dims.enc_layers},
                       text_config={"num_hidden_layers": dims.dec_layers})
    model = GlmAsrForConditionalGeneration(cfg)
    model.load_state_dict(sd, strict=True, assign=True)
    return model.eval()


def gen_model(tag: str, dims: ModelDims, cases, seed=0):
    """cases: list of (kind, n_samples, audio_seed, max_new_tokens)."""
    sd = synthetic_state_dict(dims, seed=seed)
    model = build_hf(dims, sd)
    out = {"dims": np.array([dims.enc_layers, dims.dec_layers, seed])}
    for ci, (kind, n, aseed, G) in enumerate(cases):
        xp = mo.prestep(mo.synth_audio(kind, n, aseed))
        feat, mask = hf_features(xp)
        n_audio = mo.n_audio_tokens(n)
        ids = synthetic_prompt_ids(n_audio)
        with torch.no_grad():
            enc = model.audio_tower(feat[None]).last_hidden_state[0]
            ae = model.get_audio_features(feat[None], mask[None], return_dict=True).pooler_output
            kw = dict(input_ids=torch.tensor([ids]), input_features=feat[None], input_features_mask=mask[None],
                      attention_mask=torch.ones(1, len(ids), dtype=torch.long))
            logits = model(**kw, logits_to_keep=1).logits[0, -1].float()
            gen = model.generate(**kw, max_new_tokens=G, do_sample=False, output_scores=True,
                                 return_dict_in_generate=True)
        new = gen.sequences[0, len(ids):].numpy()
        margins = []
        for s in gen.scores:
            t2 = torch.topk(s[0].float(), 2).values
            margins.append(float(t2[0] - t2[1]))
        p = f"c{ci}_"
        out[p + "case"] = np.array([n, aseed, G, n_audio])
        out[p + "kind"] = np.array(kind)
        out[p + "enc_out_sub"] = enc[::25].numpy().copy()            # [60, 1280]
        out[p + "enc_out_rms"] = np.array(float(enc.pow(2).mean().sqrt()))
        out[p + "audio_embeds_sub"] = ae[::5].numpy().copy()
        out[p + "first_logits"] = logits.numpy()
        out[p + "new_ids"] = new
        out[p + "margins"] = np.array(margins, dtype=np.float32)
        print(tag, kind, n, "ids", new[:12], "distinct", len(set(new.tolist())), "min margin", min(margins))
    np.savez_compressed(os.path.join(OUT, f"model_{tag}.npz"), **out)
    print(f"wrote model_{tag}.npz")


def short_window_T(n_samples: int) -> int:
    """Encoder positions of the opt-in short window (include/sonic_b200.h SONIC_FLAG_SHORT_WINDOW)."""
    f = mo.n_valid_frames(n_samples)
    return min(1500, ((f + 1) // 2 + 7) // 8 * 8)


def gen_short_window(seed=0):
    """The opt-in streaming-encoder mode has its own reference: the SAME HF classes fed input_features[:, :, :2T]."""
    dims = ModelDims(enc_layers=2, dec_layers=2)
    sd = synthetic_state_dict(dims, seed=seed)
    model = build_hf(dims, sd)
    out = {}
    for ci, (kind, n, aseed, G) in enumerate([("noise", 20480, 3, 15), ("speech", 20480, 7, 15), ("speech", 48000, 9, 12)]):
        xp = mo.prestep(mo.synth_audio(kind, n, aseed))
        feat, mask = hf_features(xp)
        T = short_window_T(n)
        feat, mask = feat[:, :2 * T].contiguous(), mask[:2 * T].contiguous()
        n_audio = mo.n_audio_tokens(n)
        ids = synthetic_prompt_ids(n_audio)
        with torch.no_grad():
            enc = model.audio_tower(feat[None]).last_hidden_state[0]
            ae = model.get_audio_features(feat[None], mask[None], return_dict=True).pooler_output
            kw = dict(input_ids=torch.tensor([ids]), input_features=feat[None], input_features_mask=mask[None],
                      attention_mask=torch.ones(1, len(ids), dtype=torch.long))
            gen = model.generate(**kw, max_new_tokens=G, do_sample=False, output_scores=True, return_dict_in_generate=True)
        new = gen.sequences[0, len(ids):].numpy()
        margins = [float((lambda t2: t2[0] - t2[1])(torch.topk(s[0].float(), 2).values)) for s in gen.scores]
        p = f"c{ci}_"
        out[p + "case"] = np.array([n, aseed, G, n_audio, T])
        out[p + "kind"] = np.array(kind)
        out[p + "enc_out_sub"] = enc[::4].numpy().copy()
        out[p + "audio_embeds"] = ae.numpy().copy()
        out[p + "new_ids"] = new
        out[p + "margins"] = np.array(margins, dtype=np.float32)
        print("short", kind, n, "T", T, "enc", tuple(enc.shape), "ae", tuple(ae.shape), "ids", new[:10], "min margin", min(margins))
    np.savez_compressed(os.path.join(OUT, "short_window_tiny.npz"), **out)
    print("wrote short_window_tiny.npz")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    what = sys.argv[1:] or ["mel", "tiny", "short", "full"]
    if "mel" in what:
        gen_mel()
    if "tiny" in what:
        gen_model("tiny", ModelDims(enc_layers=2, dec_layers=2),
                  [("speech", 320000, 1, 32), ("noise", 20480, 3, 15), ("speech", 163840, 11, 24)])
    if "short" in what:
        gen_short_window()
    if "full" in what:
        gen_model("full", ModelDims(), [("speech", 320000, 1, 128), ("noise", 20480, 3, 15)])
