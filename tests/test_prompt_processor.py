"""CPU tests of the prompt path against the genuine ``GlmAsrProcessor`` (transformers 5.5.0) on a locally built tokenizer
(tests/stub_checkpoint.py): ``PromptBuilder`` must produce exactly the ``input_ids`` that
``processor.apply_chat_template(tokenize=True)`` — the call of /root/reference/backend/asr.py:393-399 — produces, for every
audio length and hotword list; plus the reference's hotword cleaning rules (asr.py:303-333)."""
import numpy as np
import pytest

from sonicscribe_b200.engine import num_audio_tokens
from sonicscribe_b200.prompt import PromptBuilder, clean_hotwords, format_hotwords_prompt, instruction_text, synthetic_prompt_ids
from tests.stub_checkpoint import build_processor


@pytest.fixture(scope="module")
def processor():
    return build_processor()


def _hf_ids(processor, x, text):
    msgs = [{"role": "user", "content": [{"type": "audio", "audio": x}, {"type": "text", "text": text}]}]
    out = processor.apply_chat_template(msgs, tokenize=True, add_generation_prompt=True, return_dict=True, return_tensors="pt")
    return out["input_ids"][0].tolist(), int(out["input_features_mask"].sum())


@pytest.mark.parametrize("n", [1600, 20480, 163840, 319963, 320000, 480000])
@pytest.mark.parametrize("hotwords", [None, ["Foo", "bar"], ["Kubernetes", " B200 ", "sonic", "scribe", "foo", "", None]])
def test_prompt_ids_equal_processor_ids(processor, n, hotwords):
    x = (np.random.default_rng(n).standard_normal(n) * 0.1).astype(np.float32)
    ref_ids, frames = _hf_ids(processor, x, instruction_text(hotwords))
    pb = PromptBuilder(processor)
    got = pb.build(num_audio_tokens(n), hotwords)
    assert frames == -(-n // 160)
    assert got == ref_ids
    assert got.count(pb.audio_token_id) == num_audio_tokens(n)
    assert pb.build(num_audio_tokens(n), hotwords) == ref_ids            # cached template, same result


def test_audio_token_id_comes_from_the_processor(processor):
    pb = PromptBuilder(processor)
    assert pb.audio_token_id == processor.audio_token_id == 59260
    assert PromptBuilder(None).audio_token_id == 59260


def test_tokenizer_without_template_falls_back_to_processor(processor):
    """A tokenizer that carries no chat template of its own raises in apply_chat_template; the processor's template
    (chat_template.jinja of the checkpoint) must be used."""
    assert getattr(processor.tokenizer, "chat_template", None) in (None, "")
    pb = PromptBuilder(processor)
    pre, post = pb._template_ids(instruction_text(None))
    assert 59260 not in pre and 59260 not in post and len(pre) >= 2 and len(post) >= 2


def test_hotword_cleaning_matches_reference_rules():
    """asr.py:317-328: set() over the RAW strings, then strip/lower, then the first 10 — no second de-duplication."""
    assert format_hotwords_prompt(None) == "" and format_hotwords_prompt([]) == "" and format_hotwords_prompt(["", "  "]) == ""
    assert clean_hotwords(["Foo", "Foo", "bar"]) == ["foo", "bar"]
    assert clean_hotwords([" Foo ", "foo", "", "Bar"]) == ["foo", "foo", "bar"]          # distinct raw strings both survive
    assert clean_hotwords([1, None, "x"]) == ["x"]
    assert format_hotwords_prompt(["A", "b"]) == '. Pay special attention to these important terms: "a", "b"'
    assert len(clean_hotwords([f"w{i}" for i in range(20)])) == 10
    assert instruction_text(["a"]) == 'Please transcribe this audio into text. Pay special attention to these important terms: "a"'


def test_prompt_longer_than_capacity_trims_hotwords(processor, caplog):
    many = [f"term{i} extra words here" for i in range(10)]
    full = PromptBuilder(processor).build(375, many)
    pb = PromptBuilder(processor, max_prompt=len(full) - 12)
    with caplog.at_level("WARNING"):
        ids = pb.build(375, many)
    assert len(ids) <= len(full) - 12 and ids.count(59260) == 375
    assert any("hotwords" in r.message for r in caplog.records)
    with pytest.raises(ValueError):
        PromptBuilder(processor, max_prompt=300).build(375, None)
    # default capacity of the drop-in class holds the worst case: 30 s of audio + ten multi-token hotwords
    from sonicscribe_b200.asr import DEFAULT_MAX_PROMPT
    assert len(full) <= DEFAULT_MAX_PROMPT


def test_synthetic_prompt_without_processor():
    ids = synthetic_prompt_ids(250)
    assert len(ids) == 270 and ids.count(59260) == 250
    assert synthetic_prompt_ids(16, ["x"]) != synthetic_prompt_ids(16)
    assert PromptBuilder(None).build(16, ["x"]) == synthetic_prompt_ids(16, ["x"])


def test_batch_decode_round_trip(processor):
    ids = processor.tokenizer("please transcribe this audio")["input_ids"]
    text = processor.batch_decode([ids + [59246]], skip_special_tokens=True)[0].strip()
    assert text == "please transcribe this audio"
