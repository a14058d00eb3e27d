"""The CPU oracle (oracle/) against fixtures produced by the real HF implementation (tests/golden/gen_golden.py)
and against the bootstrap checksums recorded in SURVEY.md §8(c).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import mel_oracle as mo
from oracle import model_oracle as ora
from sonicscribe_b200.weights import ModelDims, synthetic_state_dict
from tests.golden.gen_golden import FRAME_STRIDE, MEL_CASES, synthetic_prompt_ids


def test_filter_bank_shape_and_sparsity():
    fb = mo.mel_filter_bank()
    assert fb.shape == (201, 128)
    assert int((fb.astype(np.float32) != 0).sum()) == 394
    start, count, w = mo.sparse_mel_taps()
    assert count.max() <= 9
    dense = np.zeros((201, 128), np.float32)
    for m in range(128):
        dense[start[m]:start[m] + count[m], m] = w[m, :count[m]]
    assert np.array_equal(dense, fb.astype(np.float32))


def test_token_count_formula():
    # 1 s->12, 5 s->62, 10 s->125, 20 s->250, 30 s->375
    for sec, n in [(1, 12), (5, 62), (10, 125), (20, 250), (30, 375)]:
        assert mo.n_audio_tokens(sec * 16000) == n
    assert mo.n_audio_tokens(20480) == 16
    assert mo.n_valid_frames(319963) == 2000


def test_survey_bootstrap_checksums():
    """torch.randn(320000, seed 0)*0.1 clamp, raw (no pre-step)."""
    w = (torch.randn(320000, generator=torch.Generator().manual_seed(0)) * 0.1).clamp(-1, 1).numpy()
    m, mask = mo.log_mel(w)
    assert m.shape == (128, 3000) and int(mask.sum()) == 2000
    assert abs(float(m.astype(np.float64).sum()) - 18581.900138) < 0.5
    for (i, j, v) in [(0, 0, 0.4871106), (64, 1000, 0.6216496), (127, 1999, 0.5908269), (0, 2000, 0.2420571)]:
        assert abs(float(m[i, j]) - v) < 1e-4
    assert abs(float(m.max()) - 0.9489320) < 1e-4 and abs(float(m.min()) + 1.0510681) < 1e-4


@pytest.mark.parametrize("ci", range(len(MEL_CASES)))
def test_mel_oracle_vs_hf_golden(golden_dir, ci):
    g = np.load(os.path.join(golden_dir, "mel_cases.npz"))
    kind, n, seed, pre = MEL_CASES[ci]
    x = mo.synth_audio(kind, n, seed)
    xp = mo.prestep(x) if pre else x
    m, mask = mo.log_mel(xp)
    ref = g[f"c{ci}_sub"]
    stats = g[f"c{ci}_stats"]
    assert int(mask.sum()) == int(stats[3])
    d = np.abs(m[:, ::FRAME_STRIDE] - ref)
    # HF computes the STFT in fp32; the float64 restatement agrees to ~2e-5 max-abs on every case
    assert d.max() < 5e-5, (kind, n, float(d.max()))
    assert abs(float(m.max()) - stats[1]) < 1e-4


def test_prestep_semantics():
    x = np.array([[0.5, -0.25, 0.1], [9, 9, 9]], np.float32)
    y = mo.prestep(x)
    assert y.shape == (3,)
    assert y[0] == np.float32(32767.0 / 32768.0) and y[1] == np.float32(np.rint(-0.5 * 32767) / 32768)
    z = mo.prestep(np.zeros(100, np.float32) + 1e-7)       # below the 1e-6 gate: not normalised
    assert np.all(z == 0)


def _check_model(golden_dir, tag, dims, cases_idx):
    g = np.load(os.path.join(golden_dir, f"model_{tag}.npz"))
    sd = synthetic_state_dict(dims, seed=int(g["dims"][2]))
    cfg = ora.OracleConfig(enc_layers=dims.enc_layers, dec_layers=dims.dec_layers)
    for ci in cases_idx:
        n, aseed, G, n_audio = [int(v) for v in g[f"c{ci}_case"]]
        kind = str(g[f"c{ci}_kind"])
        mel, _ = mo.log_mel(mo.prestep(mo.synth_audio(kind, n, aseed)))
        assert mo.n_audio_tokens(n) == n_audio
        probes = {}
        new, margins, fl = ora.generate_greedy(sd, cfg, torch.from_numpy(mel), n_audio, synthetic_prompt_ids(n_audio), G,
                                               probes=probes)
        enc = probes["enc_out"][::25].numpy()
        assert np.abs(enc - g[f"c{ci}_enc_out_sub"]).max() < 2e-3
        ae = probes["audio_embeds"][::5].numpy()
        assert np.abs(ae - g[f"c{ci}_audio_embeds_sub"]).max() < 2e-3
        assert np.abs(fl.numpy() - g[f"c{ci}_first_logits"]).max() < 5e-3
        assert new == g[f"c{ci}_new_ids"].tolist()


def test_model_oracle_vs_hf_golden_tiny(golden_dir):
    _check_model(golden_dir, "tiny", ModelDims(enc_layers=2, dec_layers=2), [0, 1, 2])


@pytest.mark.slow
def test_model_oracle_vs_hf_golden_full_short(golden_dir):
    # full-size model, the 1.28 s / 15-token interim case (~40 s on 8 cores)
    _check_model(golden_dir, "full", ModelDims(), [1])


def test_short_window_oracle_vs_hf_golden(golden_dir):
    """The opt-in streaming-encoder mode (SONIC_FLAG_SHORT_WINDOW) has its own reference: the HF classes fed the truncated
    features input_features[:, :, :2T].  The oracle, given the same truncated features, must reproduce them."""
    from tests.golden.gen_golden import short_window_T
    g = np.load(os.path.join(golden_dir, "short_window_tiny.npz"))
    dims = ModelDims(enc_layers=2, dec_layers=2)
    sd = synthetic_state_dict(dims, seed=0)
    cfg = ora.OracleConfig(enc_layers=2, dec_layers=2)
    for ci in range(3):
        n, aseed, G, n_audio, T = [int(v) for v in g[f"c{ci}_case"]]
        assert T == short_window_T(n) and T % 8 == 0 and T // 4 >= n_audio
        mel, _ = mo.log_mel(mo.prestep(mo.synth_audio(str(g[f"c{ci}_kind"]), n, aseed)))
        probes = {}
        new, margins, _ = ora.generate_greedy(sd, cfg, torch.from_numpy(mel[:, :2 * T].copy()), n_audio, synthetic_prompt_ids(n_audio), G,
                                              probes=probes)
        assert probes["enc_out"].shape[0] == T
        assert np.abs(probes["enc_out"][::4].numpy() - g[f"c{ci}_enc_out_sub"]).max() < 2e-3
        assert np.abs(probes["audio_embeds"].numpy() - g[f"c{ci}_audio_embeds"]).max() < 2e-3
        assert new == g[f"c{ci}_new_ids"].tolist()


def test_streamed_synthetic_checkpoint_equals_dict():
    """bench.py streams the synthetic checkpoint tensor by tensor; same names, order and values as the dict form."""
    import torch
    from sonicscribe_b200.weights import ModelDims, iter_synthetic_tensors, synthetic_state_dict
    dims = ModelDims(enc_layers=1, dec_layers=1)
    sd = synthetic_state_dict(dims, seed=3)
    names = []
    for name, t in iter_synthetic_tensors(dims, seed=3):
        names.append(name)
        if t.numel() < 4_000_000:
            assert torch.equal(t, sd[name]), name
    assert names == list(sd.keys())
