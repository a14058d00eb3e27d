"""Prompt construction for the GLM-ASR chat request (host-side integer bookkeeping, kept off the GPU).

Mirrors /root/reference/backend/asr.py:303-333,375-399 (instruction text + hotword clause) and
transformers/models/glmasr/processing_glmasr.py:97-111 (expansion of the audio placeholder to n_audio copies).
The chat template and tokenizer ship with the checkpoint; when none is available (no network in the build
container) a fixed synthetic prefix/suffix is used so parity is expressed in token ids.
"""
from __future__ import annotations

import logging
import threading
import zlib
from typing import List, Optional, Sequence

logger = logging.getLogger("sonicscribe_b200.prompt")

AUDIO_TOKEN_ID = 59260
BASE_INSTRUCTION = "Please transcribe this audio into text"
SYNTH_PREFIX = tuple(range(100, 108))
SYNTH_SUFFIX = tuple(range(200, 212))


def clean_hotwords(hotwords: Optional[Sequence[str]], max_hotwords: int = 10) -> List[str]:
    """The cleaning rules of ASRModel._format_hotwords_prompt (asr.py:317-328): de-duplicate the RAW strings (the
    reference uses ``set(hotwords)``), drop empties / non-strings, strip + lower, keep the first <=10.  ``set`` order is
    arbitrary in the reference; order of first appearance is used here so the prompt is deterministic.  Like the
    reference, two spellings that only differ in case or padding ("Foo", " foo") both survive."""
    if not hotwords:
        return []
    seen, cleaned = set(), []
    for hw in hotwords:
        try:
            if hw in seen:
                continue
            seen.add(hw)
        except TypeError:            # unhashable entries cannot be in the reference's set either (it would raise); skip
            continue
        if hw and isinstance(hw, str) and hw.strip():
            cleaned.append(hw.strip().lower())
    return cleaned[:max_hotwords]


def format_hotwords_prompt(hotwords: Optional[Sequence[str]], max_hotwords: int = 10) -> str:
    cleaned = clean_hotwords(hotwords, max_hotwords)
    if not cleaned:
        return ""
    return ". Pay special attention to these important terms: " + ", ".join(f'"{w}"' for w in cleaned)


def instruction_text(hotwords=None, max_hotwords: int = 10) -> str:
    return BASE_INSTRUCTION + format_hotwords_prompt(hotwords, max_hotwords)


def synthetic_prompt_ids(n_audio: int, hotwords=None, max_hotwords: int = 10) -> List[int]:
    """prefix(8) ++ [59260]*n_audio ++ suffix(12) (+ 2 pseudo-tokens per hotword so hotwords change the prompt)."""
    extra = []
    for w in clean_hotwords(hotwords, max_hotwords):
        hsh = zlib.crc32(w.encode())
        extra += [1000 + hsh % 50000, 1000 + (hsh >> 8) % 50000]
    return list(SYNTH_PREFIX) + [AUDIO_TOKEN_ID] * n_audio + list(SYNTH_SUFFIX) + extra


class PromptBuilder:
    """Caches the tokenised chat template per instruction text; only the audio-token run length varies per call, so the
    jinja render + tokenizer pass of asr.py:393-399 runs once per distinct hotword list instead of once per segment.

    ``max_prompt`` is the engine's per-segment prompt capacity: a hotword clause that would push the prompt past it is
    shortened one term at a time (with a warning) instead of failing the segment — the reference has no such limit."""

    def __init__(self, processor=None, max_prompt: Optional[int] = None):
        self.processor = processor
        self.max_prompt = max_prompt
        self.audio_token_id = int(getattr(processor, "audio_token_id", AUDIO_TOKEN_ID) or AUDIO_TOKEN_ID) if processor is not None else AUDIO_TOKEN_ID
        self._cache = {}
        self._lock = threading.Lock()

    # -- tokenizer-backed path ------------------------------------------------------------------------------------------
    def _render(self, text: str) -> str:
        messages = [{"role": "user", "content": [{"type": "audio", "url": "placeholder.wav"}, {"type": "text", "text": text}]}]
        last = None
        for owner in (self.processor, getattr(self.processor, "tokenizer", None)):
            if owner is None or not hasattr(owner, "apply_chat_template"):
                continue
            try:        # a tokenizer without its own chat_template raises; the processor carries the checkpoint's template
                return owner.apply_chat_template(messages, tokenize=False, add_generation_prompt=True)
            except Exception as e:   # noqa: BLE001 - fall through to the other owner
                last = e
        raise RuntimeError(f"could not render the chat template: {last}")

    def _template_ids(self, text: str):
        with self._lock:
            hit = self._cache.get(text)
        if hit is not None:
            return hit
        rendered = self._render(text)
        if isinstance(rendered, (list, tuple)):
            rendered = rendered[0]
        ids = list(self.processor.tokenizer(rendered)["input_ids"])     # same tokenizer defaults as GlmAsrProcessor.__call__
        if ids and isinstance(ids[0], (list, tuple)):
            ids = list(ids[0])
        if ids.count(self.audio_token_id) != 1:
            raise RuntimeError("chat template did not yield exactly one audio placeholder token")
        k = ids.index(self.audio_token_id)
        out = (tuple(ids[:k]), tuple(ids[k + 1:]))
        with self._lock:
            self._cache[text] = out
        return out

    def _build(self, n_audio: int, hotwords, max_hotwords: int) -> List[int]:
        if self.processor is None:
            return synthetic_prompt_ids(n_audio, hotwords, max_hotwords)
        pre, post = self._template_ids(instruction_text(hotwords, max_hotwords))
        return list(pre) + [self.audio_token_id] * n_audio + list(post)

    def build(self, n_audio: int, hotwords=None) -> List[int]:
        ids = self._build(n_audio, hotwords, 10)
        if self.max_prompt is None or len(ids) <= self.max_prompt:
            return ids
        n_terms = len(clean_hotwords(hotwords))
        for keep in range(n_terms - 1, -1, -1):
            ids = self._build(n_audio, hotwords, keep) if keep else self._build(n_audio, None, 10)
            if len(ids) <= self.max_prompt:
                logger.warning("prompt of %d audio tokens + %d hotwords exceeds max_prompt=%d; kept the first %d hotwords",
                               n_audio, n_terms, self.max_prompt, keep)
                return ids
        raise ValueError(f"prompt of {len(ids)} tokens exceeds max_prompt={self.max_prompt} even without hotwords")
