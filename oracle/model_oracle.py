"""CPU oracle for the GLM-ASR-Nano-2512 forward pass and greedy decode.

TEST INFRASTRUCTURE ONLY (see oracle/mel_oracle.py header for the import rule).

The reference repo (``backend/asr.py:407-422``) delegates all of this to the third-party
``transformers`` package (unpinned in ``backend/requirements.txt:12``; pinned here to the
container's transformers 5.5.0 / torch 2.11.0).  This file is a plain, functional torch-CPU
restatement of exactly the modules the reference call reaches:

* ``GlmAsrEncoder.forward``            transformers/models/glmasr/modeling_glmasr.py:316-330
* ``GlmAsrAttention`` / RoPE           modeling_glmasr.py:45-109,156-225
* ``GlmAsrMLP`` / ``GlmAsrEncoderLayer``   modeling_glmasr.py:228-274
* ``get_audio_features`` + projector   modeling_glmasr.py:333-349,394-426
* embed + masked_scatter               modeling_glmasr.py:473-483
* ``LlamaDecoderLayer`` etc.           transformers/models/llama/modeling_llama.py:53-67,73-168,171-184,225-332
* greedy loop                          transformers/generation/utils.py:2743-2809

Parity status: PINNED — ``tests/golden/gen_golden.py`` runs the real HF classes in this container on
seeded weights and stores their outputs (hidden-state probes, logits, greedy ids) in
``tests/golden/model_*.npz``; ``tests/test_oracle_golden.py`` checks this restatement against them.

Weights are a flat ``dict[str, Tensor]`` with HF state-dict names.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F

AUDIO_TOKEN_ID = 59260
EOS_IDS = (59246, 59253, 59255)


@dataclass
class OracleConfig:
    enc_layers: int = 32
    dec_layers: int = 28
    enc_hidden: int = 1280
    enc_heads: int = 20
    enc_inter: int = 5120
    n_mels: int = 128
    dec_hidden: int = 2048
    dec_heads: int = 16
    dec_kv_heads: int = 4
    dec_inter: int = 6144
    vocab: int = 59264
    rope_theta: float = 10000.0
    rms_eps: float = 1e-5
    ln_eps: float = 1e-5


def _rope_tables(positions: torch.Tensor, rot_dim: int, theta: float):
    """cos/sin [P, rot_dim] in fp32 (modeling_glmasr.py:96-109 / modeling_llama.py:119-136)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, rot_dim, 2, dtype=torch.int64).float() / rot_dim))
    freqs = positions.float()[:, None] * inv_freq[None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def _apply_rope(x, cos, sin):
    """x: [H, P, D]; cos/sin: [P, R] with R<=D; rotates the first R dims (modeling_glmasr.py:156-171)."""
    r = cos.shape[-1]
    xr, xp = x[..., :r], x[..., r:]
    xr = xr * cos[None] + _rotate_half(xr) * sin[None]
    return torch.cat((xr, xp), dim=-1)


# ----------------------------------------------------------------------------------------------
# encoder + adapter
# ----------------------------------------------------------------------------------------------
def encoder_forward(w: dict, cfg: OracleConfig, mel: torch.Tensor, probes: dict | None = None) -> torch.Tensor:
    """mel: [128, 3000] -> last_hidden_state [1500, 1280]  (modeling_glmasr.py:316-330)."""
    dt = w["audio_tower.conv1.weight"].dtype
    p = "audio_tower."
    x = mel.to(dt)[None]
    x = F.gelu(F.conv1d(x, w[p + "conv1.weight"], w[p + "conv1.bias"], padding=1))
    x = F.gelu(F.conv1d(x, w[p + "conv2.weight"], w[p + "conv2.bias"], stride=2, padding=1))
    h = x[0].transpose(0, 1)                                             # [T, C]
    if probes is not None:
        probes["conv_out"] = h.clone()
    T = h.shape[0]
    hd = cfg.enc_hidden // cfg.enc_heads
    cos, sin = _rope_tables(torch.arange(T), hd // 2, cfg.rope_theta)    # partial_rotary_factor 0.5
    cos, sin = cos.to(dt), sin.to(dt)
    for i in range(cfg.enc_layers):
        lp = f"{p}layers.{i}."
        u = F.layer_norm(h, (cfg.enc_hidden,), w[lp + "input_layernorm.weight"], w[lp + "input_layernorm.bias"], cfg.ln_eps)
        q = F.linear(u, w[lp + "self_attn.q_proj.weight"], w[lp + "self_attn.q_proj.bias"])
        k = F.linear(u, w[lp + "self_attn.k_proj.weight"])
        v = F.linear(u, w[lp + "self_attn.v_proj.weight"], w[lp + "self_attn.v_proj.bias"])
        q = q.view(T, cfg.enc_heads, hd).transpose(0, 1)
        k = k.view(T, cfg.enc_heads, hd).transpose(0, 1)
        v = v.view(T, cfg.enc_heads, hd).transpose(0, 1)
        q, k = _apply_rope(q, cos, sin), _apply_rope(k, cos, sin)
        s = (q @ k.transpose(1, 2)) * (hd ** -0.5)                       # no mask, non-causal (:217)
        a = torch.softmax(s.float(), dim=-1).to(dt) @ v
        a = a.transpose(0, 1).reshape(T, cfg.enc_hidden)
        h = h + F.linear(a, w[lp + "self_attn.o_proj.weight"], w[lp + "self_attn.o_proj.bias"])
        u = F.layer_norm(h, (cfg.enc_hidden,), w[lp + "post_attention_layernorm.weight"], w[lp + "post_attention_layernorm.bias"], cfg.ln_eps)
        u = F.gelu(F.linear(u, w[lp + "mlp.fc1.weight"], w[lp + "mlp.fc1.bias"]))
        h = h + F.linear(u, w[lp + "mlp.fc2.weight"], w[lp + "mlp.fc2.bias"])
        if probes is not None and i == 0:
            probes["enc_layer0"] = h.clone()
    h = F.layer_norm(h, (cfg.enc_hidden,), w[p + "norm.weight"], w[p + "norm.bias"], cfg.ln_eps)
    if probes is not None:
        probes["enc_out"] = h.clone()
    return h


def adapter_forward(w: dict, cfg: OracleConfig, enc_out: torch.Tensor, n_audio: int) -> torch.Tensor:
    """[1500,1280] -> [n_audio, 2048]  (modeling_glmasr.py:410-424, 333-349)."""
    z = enc_out.reshape(-1, cfg.enc_inter)
    z = F.gelu(F.linear(z, w["multi_modal_projector.linear_1.weight"], w["multi_modal_projector.linear_1.bias"]))
    z = F.linear(z, w["multi_modal_projector.linear_2.weight"], w["multi_modal_projector.linear_2.bias"])
    return z[:n_audio]


# ----------------------------------------------------------------------------------------------
# decoder
# ----------------------------------------------------------------------------------------------
def _rmsnorm(x, weight, eps):
    dt = x.dtype
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return weight * xf.to(dt)


class KVCache:
    def __init__(self, n_layers):
        self.k = [None] * n_layers
        self.v = [None] * n_layers

    def append(self, i, k, v):
        self.k[i] = k if self.k[i] is None else torch.cat((self.k[i], k), dim=1)
        self.v[i] = v if self.v[i] is None else torch.cat((self.v[i], v), dim=1)
        return self.k[i], self.v[i]


def decoder_forward(w: dict, cfg: OracleConfig, x: torch.Tensor, pos0: int, cache: KVCache,
                    probes: dict | None = None) -> torch.Tensor:
    """x: [S, 2048] input embeddings at positions pos0..pos0+S-1.  Returns logits [vocab] of the LAST position
    (logits_to_keep=1, generation/utils.py:2487-2491)."""
    p = "language_model.model."
    S = x.shape[0]
    dt = x.dtype
    hd = cfg.dec_hidden // cfg.dec_heads
    rep = cfg.dec_heads // cfg.dec_kv_heads
    cos, sin = _rope_tables(torch.arange(pos0, pos0 + S), hd, cfg.rope_theta)
    cos, sin = cos.to(dt), sin.to(dt)
    for i in range(cfg.dec_layers):
        lp = f"{p}layers.{i}."
        u = _rmsnorm(x, w[lp + "input_layernorm.weight"], cfg.rms_eps)
        q = F.linear(u, w[lp + "self_attn.q_proj.weight"]).view(S, cfg.dec_heads, hd).transpose(0, 1)
        k = F.linear(u, w[lp + "self_attn.k_proj.weight"]).view(S, cfg.dec_kv_heads, hd).transpose(0, 1)
        v = F.linear(u, w[lp + "self_attn.v_proj.weight"]).view(S, cfg.dec_kv_heads, hd).transpose(0, 1)
        q, k = _apply_rope(q, cos, sin), _apply_rope(k, cos, sin)
        kk, vv = cache.append(i, k, v)                                    # [KV, ctx, hd]
        ctx = kk.shape[1]
        kq = kk.repeat_interleave(rep, dim=0)
        vq = vv.repeat_interleave(rep, dim=0)
        s = (q @ kq.transpose(1, 2)) * (hd ** -0.5)                       # [H, S, ctx]
        if S > 1:
            qpos = torch.arange(pos0, pos0 + S)[:, None]
            kpos = torch.arange(ctx)[None, :]
            s = s.masked_fill(kpos > qpos, float("-inf"))
        a = torch.softmax(s.float(), dim=-1).to(dt) @ vq
        a = a.transpose(0, 1).reshape(S, cfg.dec_hidden)
        x = x + F.linear(a, w[lp + "self_attn.o_proj.weight"])
        u = _rmsnorm(x, w[lp + "post_attention_layernorm.weight"], cfg.rms_eps)
        g = F.silu(F.linear(u, w[lp + "mlp.gate_proj.weight"])) * F.linear(u, w[lp + "mlp.up_proj.weight"])
        x = x + F.linear(g, w[lp + "mlp.down_proj.weight"])
        if probes is not None and i == 0 and "dec_layer0" not in probes:
            probes["dec_layer0"] = x.clone()
    xl = _rmsnorm(x[-1:], w[p + "norm.weight"], cfg.rms_eps)
    if probes is not None and "dec_last_hidden" not in probes:
        probes["dec_last_hidden"] = xl[0].clone()
    return F.linear(xl, w["language_model.lm_head.weight"])[0]


def embed_merge(w: dict, ids: torch.Tensor, audio_embeds: torch.Tensor) -> torch.Tensor:
    """modeling_glmasr.py:473-483: E[ids] with rows where ids==59260 overwritten in order."""
    x = w["language_model.model.embed_tokens.weight"][ids].clone()
    m = ids == AUDIO_TOKEN_ID
    assert int(m.sum()) == audio_embeds.shape[0], (int(m.sum()), audio_embeds.shape)
    x[m] = audio_embeds.to(x.dtype)
    return x


@torch.no_grad()
def generate_greedy(w: dict, cfg: OracleConfig, mel: torch.Tensor, n_audio: int, ids, max_new_tokens: int,
                    eos_ids=EOS_IDS, probes: dict | None = None, logit_steps=()):
    """Full path: mel [128,3000] + prompt ids -> (new token ids list, top-2 margins list, first-step logits).
    ``logit_steps``: greedy steps (0 = prefill) whose fp32 logit rows are stored in ``probes["step_logits"][step]``."""
    ids = torch.as_tensor(ids, dtype=torch.long)
    enc = encoder_forward(w, cfg, mel, probes)
    ae = adapter_forward(w, cfg, enc, n_audio)
    if probes is not None:
        probes["audio_embeds"] = ae.clone()
    x = embed_merge(w, ids, ae)
    cache = KVCache(cfg.dec_layers)
    logits = decoder_forward(w, cfg, x, 0, cache, probes)
    first_logits = logits.float().clone()
    out, margins = [], []
    pos = ids.shape[0]
    for _ in range(max_new_tokens):
        lf = logits.float()
        if probes is not None and len(out) in logit_steps:
            probes.setdefault("step_logits", {})[len(out)] = lf.clone()
        top2 = torch.topk(lf, 2)
        tok = int(torch.argmax(lf))                                      # first max index on ties
        out.append(tok)
        margins.append(float(top2.values[0] - top2.values[1]))
        if tok in eos_ids or len(out) >= max_new_tokens:
            break
        x = w["language_model.model.embed_tokens.weight"][tok][None]
        logits = decoder_forward(w, cfg, x, pos, cache)
        pos += 1
    return out, margins, first_logits


# ----------------------------------------------------------------------------------------------
# INT8 weight-only oracle (north-star variant; SURVEY.md A.6 — bitsandbytes itself is absent => unpinned)
# ----------------------------------------------------------------------------------------------
def quantize_rowwise_int8(wt: torch.Tensor):
    """Per-output-row absmax int8: s = max|W_row|, q = rint(127 W / s).  Returns (q int8 [out,in], s fp32 [out])."""
    wf = wt.float()
    s = wf.abs().amax(dim=1).clamp_min(1e-30)
    q = torch.round(wf * (127.0 / s)[:, None]).clamp(-127, 127).to(torch.int8)
    return q, s


def int8_weight_only_state(w: dict) -> dict:
    """Replace every nn.Linear weight the reference quantises (backend/asr.py:173-177: all Linear except names
    containing lm_head / embed_tokens / audio_proj) by its dequantised int8 image W' = q * s / 127."""
    out = {}
    for name, t in w.items():
        is_linear = name.endswith("_proj.weight") or ".mlp.fc" in name and name.endswith("weight") \
            or "multi_modal_projector.linear_" in name and name.endswith("weight")
        if is_linear and t.dim() == 2 and "lm_head" not in name and "embed_tokens" not in name:
            q, s = quantize_rowwise_int8(t)
            out[name] = (q.float() * (s / 127.0)[:, None]).to(t.dtype)
        else:
            out[name] = t
    return out
