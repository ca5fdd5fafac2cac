"""Model geometry and special-token layout for every Whisper size the reference names.

The reference resolves a size name to a hub repo (``faster_whisper/utils.py:11-31``) and lets
CTranslate2 read the geometry from ``model.bin``.  There is no network here, so the geometry table
(OpenAI Whisper dims) lives in-tree and a model can be built from a size name
plus a weight source (synthetic, or a loader for a converted directory).

The special-token ids follow the name->id rule the reference's tokenizer wrapper relies on
(``faster_whisper/tokenizer.py:42-78``) and the golden ids in ``tests/test_tokenizer.py:96-110``.
"""

from dataclasses import dataclass, asdict
from typing import Dict, List

N_AUDIO_CTX = 1500
N_TEXT_CTX = 448
N_TIMESTAMPS = 1501  # 0.00 .. 30.00 s in 0.02 s steps


@dataclass(frozen=True)
class WhisperDims:
    name: str
    n_mels: int
    n_audio_state: int
    n_audio_head: int
    n_audio_layer: int
    n_text_state: int
    n_text_head: int
    n_text_layer: int
    n_vocab: int
    n_audio_ctx: int = N_AUDIO_CTX
    n_text_ctx: int = N_TEXT_CTX

    @property
    def is_multilingual(self) -> bool:
        return self.n_vocab >= 51865

    @property
    def num_languages(self) -> int:
        # .en vocabularies still carry the 99 language slots (ids 50258..50356)
        return 100 if self.n_vocab == 51866 else 99

    def to_dict(self) -> dict:
        return asdict(self)


def _dims(name, mels, d, heads, enc_layers, dec_layers, vocab):
    return WhisperDims(name, mels, d, heads, enc_layers, d, heads, dec_layers, vocab)


_V_EN, _V_ML, _V_V3 = 51864, 51865, 51866

MODEL_DIMS: Dict[str, WhisperDims] = {
    d.name: d
    for d in [
        _dims("tiny.en", 80, 384, 6, 4, 4, _V_EN),
        _dims("tiny", 80, 384, 6, 4, 4, _V_ML),
        _dims("base.en", 80, 512, 8, 6, 6, _V_EN),
        _dims("base", 80, 512, 8, 6, 6, _V_ML),
        _dims("small.en", 80, 768, 12, 12, 12, _V_EN),
        _dims("small", 80, 768, 12, 12, 12, _V_ML),
        _dims("medium.en", 80, 1024, 16, 24, 24, _V_EN),
        _dims("medium", 80, 1024, 16, 24, 24, _V_ML),
        _dims("large-v1", 80, 1280, 20, 32, 32, _V_ML),
        _dims("large-v2", 80, 1280, 20, 32, 32, _V_ML),
        _dims("large-v3", 128, 1280, 20, 32, 32, _V_V3),
        _dims("large", 128, 1280, 20, 32, 32, _V_V3),
        _dims("distil-large-v2", 80, 1280, 20, 32, 2, _V_ML),
        _dims("distil-medium.en", 80, 1024, 16, 24, 2, _V_EN),
        _dims("distil-small.en", 80, 768, 12, 12, 4, _V_EN),
        _dims("distil-large-v3", 128, 1280, 20, 32, 2, _V_V3),
        _dims("distil-large-v3.5", 128, 1280, 20, 32, 2, _V_V3),
        _dims("large-v3-turbo", 128, 1280, 20, 32, 4, _V_V3),
        _dims("turbo", 128, 1280, 20, 32, 4, _V_V3),
    ]
}


@dataclass(frozen=True)
class SpecialTokens:
    """Ids of the control tokens, derived from the vocabulary size alone."""

    eot: int
    sot: int
    lang_begin: int  # id of "<|en|>"
    num_languages: int
    translate: int
    transcribe: int
    sot_lm: int
    sot_prev: int
    no_speech: int
    no_timestamps: int
    timestamp_begin: int
    n_vocab: int

    @property
    def lang_ids(self) -> List[int]:
        return list(range(self.lang_begin, self.lang_begin + self.num_languages))

    def to_dict(self) -> dict:
        return asdict(self)


def special_tokens(n_vocab: int) -> SpecialTokens:
    """Layout: <text vocab> eot sot <langs> translate transcribe sot_lm sot_prev no_speech
    no_timestamps <1501 timestamps>.  Checked against tests/test_tokenizer.py:96-110 (tiny.en)."""
    if n_vocab == _V_EN:
        eot, nlang = 50256, 99
    elif n_vocab == _V_ML:
        eot, nlang = 50257, 99
    elif n_vocab == _V_V3:
        eot, nlang = 50257, 100
    else:
        # small test vocabularies: same relative layout, 3 language slots
        nlang = 3
        eot = n_vocab - N_TIMESTAMPS - 6 - nlang - 2
        if eot < 300:
            raise ValueError(f"vocabulary of {n_vocab} is too small for the Whisper control tokens")
    sot = eot + 1
    lang_begin = sot + 1
    translate = lang_begin + nlang
    st = SpecialTokens(
        eot=eot,
        sot=sot,
        lang_begin=lang_begin,
        num_languages=nlang,
        translate=translate,
        transcribe=translate + 1,
        sot_lm=translate + 2,
        sot_prev=translate + 3,
        no_speech=translate + 4,
        no_timestamps=translate + 5,
        timestamp_begin=translate + 6,
        n_vocab=n_vocab,
    )
    assert st.timestamp_begin + N_TIMESTAMPS == n_vocab, (st, n_vocab)
    return st


# Language codes in id order (reference: faster_whisper/tokenizer.py:214-320 — a data table).
LANGUAGE_CODES = (
    "en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur "
    "hr bg lt la mi ml cy sk te fa lv bn sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si "
    "km sn yo so af oc ka be tg sd gu am yi lo uz fo ht ps tk nn mt sa lb my bo tl mg as tt haw ln ha "
    "ba jw su yue"
).split()
assert len(LANGUAGE_CODES) == 100
