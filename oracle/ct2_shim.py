"""A ``ctranslate2``-shaped module whose ``models.Whisper`` is backed by the CPU oracle.

TEST INFRASTRUCTURE ONLY.  Two uses, both on CPU in tests:
  * bound into ``sys.modules['ctranslate2']`` so the UNMODIFIED reference ``faster_whisper/transcribe.py`` runs over it;
  * monkeypatched over ``faster_whisper_b200.engine.Whisper`` so our own host layer runs without a GPU.
Comparing the two runs checks that our host layer is a drop-in for the reference's (same segments, same info).
"""
import types
import zlib
from dataclasses import dataclass, field
from typing import List

import numpy as np

from faster_whisper_b200.config import LANGUAGE_CODES, special_tokens
from oracle.whisper_oracle import WhisperOracle


class StorageView:
    def __init__(self, array=None, tensor=None):
        self.array = array
        self.tensor = tensor
        self.shape = list(array.shape if array is not None else tensor.shape)
        self._handle = tensor  # the engine wrapper checks this attribute

    @classmethod
    def from_array(cls, array):
        assert array.flags["C_CONTIGUOUS"]
        return cls(array=array)

    def numpy(self):
        return self.array if self.array is not None else self.tensor.numpy()


@dataclass
class WhisperGenerationResult:
    sequences_ids: List[List[int]]
    scores: List[float]
    no_speech_prob: float
    sequences: List[List[str]] = field(default_factory=list)


def make_whisper_class(dims, weights, calls=None):
    """Returns a class with ctranslate2.models.Whisper's constructor signature bound to one synthetic checkpoint."""
    st = special_tokens(dims.n_vocab)
    oracle = WhisperOracle(dims.to_dict(), weights, st.to_dict())

    class Whisper:
        def __init__(self, model_path="", device="auto", device_index=0, compute_type="default", intra_threads=0,
                     inter_threads=1, files=None, **kwargs):
            self.dims = dims
            self.tokens = st
            self._device_index = [device_index] if isinstance(device_index, int) else list(device_index)
            self.compute_type = compute_type

        is_multilingual = property(lambda self: dims.is_multilingual)
        n_mels = property(lambda self: dims.n_mels)
        device = property(lambda self: "cpu")
        device_index = property(lambda self: list(self._device_index))

        def encode(self, features, to_cpu=False):
            arr = features.numpy() if hasattr(features, "numpy") else np.asarray(features)
            if calls is not None:  # shape + checksum of the features: the drop-in tests compare what reaches the engine, bit for bit
                calls.append(("encode", arr.shape, zlib.crc32(np.ascontiguousarray(arr, dtype=np.float32).tobytes())))
            return StorageView(tensor=oracle.encode(arr))

        def generate(self, features, prompts, **kw):
            if calls is not None:
                calls.append(("generate", [list(p) for p in prompts], {k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()}))
            enc = features.tensor if features.tensor is not None else oracle.encode(features.array)
            kw.pop("asynchronous", None)
            res = oracle.generate(enc, prompts, **kw)
            return [WhisperGenerationResult(r.sequences_ids, r.scores, r.no_speech_prob) for r in res]

        def detect_language(self, features):
            enc = features.tensor if features.tensor is not None else oracle.encode(features.array)
            names = (LANGUAGE_CODES + ["xx1", "xx2", "xx3"])
            return [[(f"<|{names[i - st.lang_begin]}|>", p) for i, p in row] for row in oracle.detect_language(enc)]

        def align(self, features, start_sequence, text_tokens, num_frames, median_filter_width=7):
            if calls is not None:
                calls.append(("align", list(start_sequence), [list(t) for t in text_tokens], num_frames, median_filter_width))
            enc = features.tensor if features.tensor is not None else oracle.encode(features.array)
            return oracle.align(enc, start_sequence, text_tokens, num_frames, median_filter_width)

        def encode_audio(self, chunks, return_features=False):
            from oracle.whisper_oracle import log_mel, pad_or_trim

            assert all(len(c) <= 30 * 16000 for c in chunks), "the fused audio path is specified for chunks of at most 30 s"
            feats = np.stack([pad_or_trim(log_mel(c, dims.n_mels)[:, :-1]) for c in chunks])
            sv = self.encode(StorageView.from_array(np.ascontiguousarray(feats)))
            return (sv, feats) if return_features else sv

    return Whisper, oracle


def make_module(dims, weights, calls=None):
    whisper_cls, oracle = make_whisper_class(dims, weights, calls)
    mod = types.ModuleType("ctranslate2")
    mod.models = types.ModuleType("ctranslate2.models")
    mod.models.Whisper = whisper_cls
    mod.models.WhisperGenerationResult = WhisperGenerationResult
    mod.StorageView = StorageView
    return mod, oracle
