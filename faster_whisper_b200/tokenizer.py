"""Thin wrapper over a Hugging Face ``tokenizers.Tokenizer`` exposing Whisper's control-token ids.

Same interface as the reference wrapper (``faster_whisper/tokenizer.py:9-211``): the hot path needs the ids
(``sot_sequence``, ``no_timestamps``, ``timestamp_begin``, ``non_speech_tokens`` for the device-side masks);
the string work stays on the host.
"""

from __future__ import annotations

import string
from functools import cached_property
from typing import List, Optional, Tuple

from .config import LANGUAGE_CODES

_TASKS = ("transcribe", "translate")
_LANGUAGE_CODES = tuple(sorted(LANGUAGE_CODES[:-1])) + ("yue",)
_NO_SPACE_LANGUAGES = frozenset({"zh", "ja", "th", "lo", "my", "yue"})


class Tokenizer:
    def __init__(self, tokenizer, multilingual: bool, task: Optional[str] = None, language: Optional[str] = None):
        self.tokenizer = tokenizer
        if not multilingual:
            self.task = self.language = None
            self.language_code = "en"
            return
        if task not in _TASKS:
            raise ValueError("'%s' is not a valid task (accepted tasks: %s)" % (task, ", ".join(_TASKS)))
        if language not in _LANGUAGE_CODES:
            raise ValueError("'%s' is not a valid language code (accepted language codes: %s)"
                             % (language, ", ".join(_LANGUAGE_CODES)))
        self.task = self._id(f"<|{task}|>")
        self.language = self._id(f"<|{language}|>")
        self.language_code = language

    def _id(self, token: str) -> Optional[int]:
        return self.tokenizer.token_to_id(token)

    # --- control tokens -------------------------------------------------------------------------------
    @cached_property
    def transcribe(self) -> int:
        return self._id("<|transcribe|>")

    @cached_property
    def translate(self) -> int:
        return self._id("<|translate|>")

    @cached_property
    def sot(self) -> int:
        return self._id("<|startoftranscript|>")

    @cached_property
    def sot_lm(self) -> int:
        return self._id("<|startoflm|>")

    @cached_property
    def sot_prev(self) -> int:
        return self._id("<|startofprev|>")

    @cached_property
    def eot(self) -> int:
        return self._id("<|endoftext|>")

    @cached_property
    def no_timestamps(self) -> int:
        return self._id("<|notimestamps|>")

    @cached_property
    def no_speech(self) -> int:
        return self._id("<|nospeech|>") or self._id("<|nocaptions|>")

    @property
    def timestamp_begin(self) -> int:
        return self.no_timestamps + 1

    @property
    def sot_sequence(self) -> List[int]:
        return [t for t in (self.sot, self.language, self.task) if t is not None]

    # --- text <-> ids -----------------------------------------------------------------------------------
    def encode(self, text: str) -> List[int]:
        return self.tokenizer.encode(text, add_special_tokens=False).ids

    def decode(self, tokens: List[int]) -> str:
        return self.tokenizer.decode([t for t in tokens if t < self.eot])

    def decode_with_timestamps(self, tokens: List[int]) -> str:
        pieces: List[str] = []
        run: List[int] = []
        for t in tokens:
            if t >= self.timestamp_begin:
                pieces.append(self.tokenizer.decode(run))
                run = []
                pieces.append(f"<|{(t - self.timestamp_begin) * 0.02:.2f}|>")
            else:
                run.append(t)
        pieces.append(self.tokenizer.decode(run))
        return "".join(pieces)

    @cached_property
    def non_speech_tokens(self) -> Tuple[int]:
        """Ids of speaker tags / annotation symbols to suppress (basic punctuation is kept).  Symbols that
        tokenise to several ids only contribute their first id when they are musical-note characters (their
        UTF-8 encodings share the first two bytes, so the first id is safe to suppress)."""
        singles = list('"#()*+/:;<=>@[\\]^_`{|}~「」『』')
        multi = "<< >> <<< >>> -- --- -( -[ (' (\" (( )) ((( ))) [[ ]] {{ }} ♪♪ ♪♪♪".split()
        notes = set("♩♪♫♬♭♮♯")
        ids = {self.encode(" -")[0], self.encode(" '")[0]}
        for sym in singles + multi + sorted(notes):
            for variant in (sym, " " + sym):
                toks = self.encode(variant)
                if len(toks) == 1 or sym in notes:
                    ids.add(toks[0])
        return tuple(sorted(ids))

    # --- word splitting (word timestamps) ------------------------------------------------------------------
    def split_to_word_tokens(self, tokens: List[int]) -> Tuple[List[str], List[List[int]]]:
        if self.language_code in _NO_SPACE_LANGUAGES:
            return self.split_tokens_on_unicode(tokens)
        return self.split_tokens_on_spaces(tokens)

    def split_tokens_on_unicode(self, tokens: List[int]) -> Tuple[List[str], List[List[int]]]:
        """Cuts wherever the accumulated ids decode to complete code points."""
        whole = self.decode_with_timestamps(tokens)
        bad = "�"
        words, groups, pending, consumed = [], [], [], 0
        for t in tokens:
            pending.append(t)
            text = self.decode_with_timestamps(pending)
            at = text.find(bad)
            if at < 0 or (consumed + at < len(whole) and whole[consumed + at] == bad):
                words.append(text)
                groups.append(pending)
                pending = []
                consumed += len(text)
        return words, groups

    def split_tokens_on_spaces(self, tokens: List[int]) -> Tuple[List[str], List[List[int]]]:
        pieces, piece_ids = self.split_tokens_on_unicode(tokens)
        words: List[str] = []
        groups: List[List[int]] = []
        for text, ids in zip(pieces, piece_ids):
            starts_word = (ids[0] >= self.eot or text.startswith(" ") or text.strip() in string.punctuation or not words)
            if starts_word:
                words.append(text)
                groups.append(list(ids))
            else:
                words[-1] += text
                groups[-1].extend(ids)
        return words, groups
