"""Checkpoint handling for the GLM-ASR-Nano-2512 path.

Weight names/shapes are the HF state-dict contract the reference loads with
``AutoModel.from_pretrained`` (/root/reference/backend/asr.py:137-140).

* ``ModelDims``            — architecture constants (transformers/models/glmasr/configuration_glmasr.py:34-110).
* ``synthetic_state_dict`` — seeded random checkpoint (no network => no real checkpoint); every tensor has its own
                             generator so any subset can be regenerated bit-identically on any machine.
* ``load_checkpoint_dir``  — read ``*.safetensors`` from a checkpoint directory into the same flat dict.
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass, asdict
from pathlib import Path

import torch


@dataclass(frozen=True)
class ModelDims:
    enc_layers: int = 32
    dec_layers: int = 28
    n_mels: int = 128
    enc_hidden: int = 1280
    enc_heads: int = 20
    enc_inter: int = 5120
    dec_hidden: int = 2048
    dec_heads: int = 16
    dec_kv_heads: int = 4
    dec_inter: int = 6144
    vocab: int = 59264
    audio_token_id: int = 59260
    eos_ids: tuple = (59246, 59253, 59255)

    def as_dict(self):
        return asdict(self)


def tensor_specs(d: ModelDims):
    """Yield (name, shape, kind) for every tensor of the checkpoint, in HF state-dict naming."""
    a = "audio_tower."
    yield a + "conv1.weight", (d.enc_hidden, d.n_mels, 3), "w"
    yield a + "conv1.bias", (d.enc_hidden,), "b"
    yield a + "conv2.weight", (d.enc_hidden, d.enc_hidden, 3), "w"
    yield a + "conv2.bias", (d.enc_hidden,), "b"
    for i in range(d.enc_layers):
        p = f"{a}layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            yield p + f"self_attn.{n}.weight", (d.enc_hidden, d.enc_hidden), "w"
            if n != "k_proj":
                yield p + f"self_attn.{n}.bias", (d.enc_hidden,), "b"
        yield p + "mlp.fc1.weight", (d.enc_inter, d.enc_hidden), "w"
        yield p + "mlp.fc1.bias", (d.enc_inter,), "b"
        yield p + "mlp.fc2.weight", (d.enc_hidden, d.enc_inter), "w"
        yield p + "mlp.fc2.bias", (d.enc_hidden,), "b"
        for n in ("input_layernorm", "post_attention_layernorm"):
            yield p + n + ".weight", (d.enc_hidden,), "g"
            yield p + n + ".bias", (d.enc_hidden,), "beta"
    yield a + "norm.weight", (d.enc_hidden,), "g"
    yield a + "norm.bias", (d.enc_hidden,), "beta"
    m = "multi_modal_projector."
    yield m + "linear_1.weight", (2 * d.dec_hidden, d.enc_inter), "w"
    yield m + "linear_1.bias", (2 * d.dec_hidden,), "b"
    yield m + "linear_2.weight", (d.dec_hidden, 2 * d.dec_hidden), "w"
    yield m + "linear_2.bias", (d.dec_hidden,), "b"
    l = "language_model.model."
    yield l + "embed_tokens.weight", (d.vocab, d.dec_hidden), "emb"
    hd = d.dec_hidden // d.dec_heads
    for i in range(d.dec_layers):
        p = f"{l}layers.{i}."
        yield p + "self_attn.q_proj.weight", (d.dec_heads * hd, d.dec_hidden), "w"
        yield p + "self_attn.k_proj.weight", (d.dec_kv_heads * hd, d.dec_hidden), "w"
        yield p + "self_attn.v_proj.weight", (d.dec_kv_heads * hd, d.dec_hidden), "w"
        yield p + "self_attn.o_proj.weight", (d.dec_hidden, d.dec_heads * hd), "w"
        yield p + "mlp.gate_proj.weight", (d.dec_inter, d.dec_hidden), "w"
        yield p + "mlp.up_proj.weight", (d.dec_inter, d.dec_hidden), "w"
        yield p + "mlp.down_proj.weight", (d.dec_hidden, d.dec_inter), "w"
        yield p + "input_layernorm.weight", (d.dec_hidden,), "g"
        yield p + "post_attention_layernorm.weight", (d.dec_hidden,), "g"
    yield l + "norm.weight", (d.dec_hidden,), "g"
    yield "language_model.lm_head.weight", (d.vocab, d.dec_hidden), "head"


# std of the seeded init per tensor kind.  Linear/conv weights follow HF's initializer_range 0.02; biases and norm
# affine terms are made non-trivial on purpose (HF would zero them) so every term of the path is exercised, and the
# embedding / lm_head scales are chosen so greedy decoding of the random model produces diverse tokens with top-2
# margins far above fp32 re-association noise (SURVEY.md §7 hard part 6).
_INIT_STD = {"w": 0.02, "b": 0.02, "g": 0.1, "beta": 0.1, "emb": 1.0, "head": 0.05}


def synthetic_tensor(name: str, shape, kind: str, seed: int) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    t = torch.randn(shape, generator=g, dtype=torch.float32) * _INIT_STD[kind]
    if kind == "g":
        t += 1.0
    return t


def synthetic_state_dict(dims: ModelDims = ModelDims(), seed: int = 0, dtype=torch.float32) -> dict:
    return {n: synthetic_tensor(n, s, k, seed).to(dtype) for n, s, k in tensor_specs(dims)}


def iter_synthetic_tensors(dims: ModelDims = ModelDims(), seed: int = 0, dtype=torch.float32):
    """The same tensors as ``synthetic_state_dict``, one (name, tensor) at a time: the full-size checkpoint is 9 GB in fp32,
    which eight benchmark ranks on one host should not each hold while their engine only needs one tensor at a time."""
    for n, s, k in tensor_specs(dims):
        yield n, synthetic_tensor(n, s, k, seed).to(dtype)


def load_checkpoint_dir(path: str) -> dict:
    """Flat state dict from a HF checkpoint directory (``model*.safetensors``)."""
    from safetensors.torch import load_file  # local import: only needed with a real checkpoint

    out = {}
    files = sorted(Path(path).glob("*.safetensors"))
    if not files:
        raise FileNotFoundError(f"no *.safetensors under {path}")
    for f in files:
        out.update(load_file(str(f)))
    # HF-native GLM-ASR checkpoints may carry a leading "model." prefix
    return {k[6:] if k.startswith("model.") and not k.startswith("model.layers") else k: v for k, v in out.items()}


def dims_from_state_dict(sd: dict) -> ModelDims:
    enc = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("audio_tower.layers."))
    dec = 1 + max(int(k.split(".")[3]) for k in sd if k.startswith("language_model.model.layers."))
    return ModelDims(enc_layers=enc, dec_layers=dec)
