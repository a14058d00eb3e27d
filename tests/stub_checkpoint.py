"""Test infrastructure: a local GLM-ASR checkpoint directory built without network access.

The real GLM-ASR-Nano-2512 checkpoint (weights + Qwen2 tokenizer + chat template) cannot be fetched here, so the
checkpoint-loading path of /root/reference/backend/asr.py:66-82,393-399,425-429 (``AutoProcessor.from_pretrained``,
``apply_chat_template``, ``batch_decode``, safetensors load) is exercised against a directory this module writes:

* ``model.safetensors``      — the seeded synthetic state dict (HF tensor names), bf16
* ``tokenizer.json`` etc.    — a 59264-entry word-level tokenizer whose special ids match the GLM-ASR config
                               (audio placeholder ``<|pad|>`` = 59260, EOS set {59246, 59253, 59255})
* ``chat_template.jinja``    — a GLM-style template with one audio placeholder per audio content entry
* ``processor_config.json``  — so that ``AutoProcessor.from_pretrained(dir)`` returns the real ``GlmAsrProcessor`` class

The processor class, feature extractor and template plumbing are the genuine transformers 5.5.0 ones; only the vocabulary
is a stand-in.
"""
from __future__ import annotations

import os

VOCAB = 59264
SPECIAL = {59260: "<|pad|>", 59246: "<|endoftext|>", 59253: "<|user|>", 59255: "<|assistant|>", 59254: "<|system|>",
           59256: "<|begin_of_audio|>", 59257: "<|end_of_audio|>"}
WORDS = ["please", "transcribe", "this", "audio", "into", "text", ".", "pay", "special", "attention", "to", "these",
         "important", "terms", ":", ",", '"', "foo", "bar", "kubernetes", "b200", "sonic", "scribe"]
CHAT_TEMPLATE = (
    "{% for m in messages %}<|{{ m['role'] }}|>\n"
    "{% for c in m['content'] %}{% if c['type'] == 'audio' %}<|begin_of_audio|><|pad|><|end_of_audio|>"
    "{% else %}{{ c['text'] }}{% endif %}{% endfor %}{% endfor %}"
    "{% if add_generation_prompt %}<|assistant|>\n{% endif %}")


def build_processor():
    from tokenizers import Tokenizer, decoders, models, normalizers, pre_tokenizers
    from transformers import PreTrainedTokenizerFast, WhisperFeatureExtractor
    from transformers.models.glmasr.processing_glmasr import GlmAsrProcessor

    vocab = {"<unk>": 0}
    for i, w in enumerate(WORDS):
        vocab[w] = 10 + i
    for i, t in SPECIAL.items():
        vocab[t] = i
    used = set(vocab.values())
    for i in range(VOCAB):
        if i not in used:
            vocab[f"w{i}"] = i
    tok = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    tok.normalizer = normalizers.Lowercase()
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.WhitespaceSplit(), pre_tokenizers.Punctuation()])
    tok.decoder = decoders.WordPiece(prefix="##", cleanup=False)       # joins tokens with single spaces
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="<unk>", pad_token="<|endoftext|>", eos_token="<|endoftext|>",
                                   additional_special_tokens=list(SPECIAL.values()))
    return GlmAsrProcessor(WhisperFeatureExtractor(feature_size=128), fast, chat_template=CHAT_TEMPLATE)


def write_checkpoint(path: str, dims=None, seed: int = 0) -> str:
    """Write processor files + ``model.safetensors`` (bf16) under ``path``; returns ``path``."""
    import torch
    from safetensors.torch import save_file

    from sonicscribe_b200.weights import ModelDims, synthetic_state_dict

    os.makedirs(path, exist_ok=True)
    build_processor().save_pretrained(path)
    dims = dims or ModelDims(enc_layers=2, dec_layers=2)
    sd = {k: v.to(torch.bfloat16).contiguous() for k, v in synthetic_state_dict(dims, seed=seed).items()}
    # what a real export also carries and the path must tolerate: a non-persistent style buffer the model never reads
    sd["language_model.model.rotary_emb.inv_freq"] = torch.arange(64, dtype=torch.float32)
    save_file(sd, os.path.join(path, "model.safetensors"))
    return path
