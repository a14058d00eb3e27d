"""Prompt construction for the GLM-ASR chat request (host-side integer bookkeeping, kept off the GPU).

Mirrors /root/reference/backend/asr.py:303-333,375-399 (instruction text + hotword clause) and
transformers/models/glmasr/processing_glmasr.py:97-111 (expansion of the audio placeholder to n_audio copies).
The chat template and tokenizer ship with the checkpoint; when none is available (no network in the build
container) a fixed synthetic prefix/suffix is used so parity is expressed in token ids.
"""
from __future__ import annotations

import zlib
from functools import lru_cache
from typing import List, Optional, Sequence

AUDIO_TOKEN_ID = 59260
BASE_INSTRUCTION = "Please transcribe this audio into text"
SYNTH_PREFIX = tuple(range(100, 108))
SYNTH_SUFFIX = tuple(range(200, 212))


def format_hotwords_prompt(hotwords: Optional[Sequence[str]], max_hotwords: int = 10) -> str:
    """Same cleaning rules as ASRModel._format_hotwords_prompt (asr.py:303-333): strip/lower, de-duplicate, first <=10.
    The reference de-duplicates through ``set`` (arbitrary order); order of first appearance is kept here so the
    prompt is deterministic."""
    if not hotwords:
        return ""
    seen, cleaned = set(), []
    for hw in hotwords:
        if hw and isinstance(hw, str) and hw.strip():
            w = hw.strip().lower()
            if w not in seen:
                seen.add(w)
                cleaned.append(w)
    if not cleaned:
        return ""
    cleaned = cleaned[:max_hotwords]
    return ". Pay special attention to these important terms: " + ", ".join(f'"{w}"' for w in cleaned)


def instruction_text(hotwords=None) -> str:
    return BASE_INSTRUCTION + format_hotwords_prompt(hotwords)


def synthetic_prompt_ids(n_audio: int, hotwords=None) -> List[int]:
    """prefix(8) ++ [59260]*n_audio ++ suffix(12) (+ 2 pseudo-tokens per hotword so hotwords change the prompt)."""
    extra = []
    clause = format_hotwords_prompt(hotwords)
    if clause:
        for w in clause.split('"')[1::2]:
            hsh = zlib.crc32(w.encode())
            extra += [1000 + hsh % 50000, 1000 + (hsh >> 8) % 50000]
    return list(SYNTH_PREFIX) + [AUDIO_TOKEN_ID] * n_audio + list(SYNTH_SUFFIX) + extra


class PromptBuilder:
    """Caches the tokenised chat template per hotword clause; only the audio-token run length varies per call."""

    def __init__(self, processor=None):
        self.processor = processor
        self._cache = {}

    def _template_ids(self, text: str):
        if text in self._cache:
            return self._cache[text]
        tok = self.processor.tokenizer
        messages = [{"role": "user", "content": [{"type": "audio", "url": "placeholder.wav"}, {"type": "text", "text": text}]}]
        rendered = tok.apply_chat_template(messages, tokenize=False, add_generation_prompt=True) \
            if hasattr(tok, "apply_chat_template") else self.processor.apply_chat_template(messages, tokenize=False, add_generation_prompt=True)
        ids = tok(rendered, add_special_tokens=False)["input_ids"]
        if ids.count(AUDIO_TOKEN_ID) != 1:
            raise RuntimeError("chat template did not yield exactly one audio placeholder token")
        k = ids.index(AUDIO_TOKEN_ID)
        self._cache[text] = (tuple(ids[:k]), tuple(ids[k + 1:]))
        return self._cache[text]

    def build(self, n_audio: int, hotwords=None) -> List[int]:
        if self.processor is None:
            return synthetic_prompt_ids(n_audio, hotwords)
        pre, post = self._template_ids(instruction_text(hotwords))
        return list(pre) + [AUDIO_TOKEN_ID] * n_audio + list(post)
