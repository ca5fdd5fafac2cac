"""-m gpu: BASELINE.json configs[0] — tiny.en geometry (d 384, 6 heads, 4+4 layers, V 51864), greedy (beam_size=1), one 30 s
synthetic chunk through ``WhisperModel.transcribe`` on the CUDA engine, token-exact against the same host code driven over
the oracle-backed engine shim (the reference's own pin for this configuration is tests/test_transcribe.py:14-59, which needs
hub weights + PyAV and cannot run offline).  Also the batched pipeline at the same geometry."""
import json

import numpy as np
import pytest

from faster_whisper_b200 import engine as our_engine
from faster_whisper_b200 import transcribe as T
from faster_whisper_b200.config import MODEL_DIMS
from faster_whisper_b200.synthetic import make_tokenizer, make_weights, synthetic_audio
from oracle import ct2_shim
from oracle import whisper_oracle as orc

pytestmark = pytest.mark.gpu

KW = dict(beam_size=1, temperature=0.0, condition_on_previous_text=False, no_speech_threshold=None, log_prob_threshold=None,
          compression_ratio_threshold=None)


@pytest.fixture(scope="module")
def tiny_en():
    dims = MODEL_DIMS["tiny.en"]
    w = make_weights(dims, seed=5)
    files = {"tokenizer.json": make_tokenizer(dims.n_vocab).to_str().encode(),
             "preprocessor_config.json": json.dumps({"feature_size": dims.n_mels}).encode()}
    gpu = T.WhisperModel("synthetic", device="cuda", files=dict(files), dims=dims, weights=w)
    calls = []
    whisper_cls, oracle = ct2_shim.make_whisper_class(dims, w, calls)
    mp = pytest.MonkeyPatch()
    mp.setattr(our_engine, "Whisper", lambda *a, **k: whisper_cls())
    mp.setattr(our_engine, "log_mel", lambda x, n_mels, padding=160, device=0: orc.log_mel(x, n_mels, padding))
    mp.setattr(our_engine, "StorageView", ct2_shim.StorageView)
    cpu = T.WhisperModel("synthetic", device="cuda", files=dict(files), dims=dims, weights=w)
    mp.undo()
    return gpu, cpu, mp, whisper_cls


def _tokens(segs):
    return [list(s.tokens) for s in segs]


def _with_shim(mp, fn):
    mp.setattr(our_engine, "log_mel", lambda x, n_mels, padding=160, device=0: orc.log_mel(x, n_mels, padding))
    mp.setattr(our_engine, "StorageView", ct2_shim.StorageView)
    try:
        return fn()
    finally:
        mp.undo()


@pytest.mark.parametrize("extra", [dict(max_new_tokens=48, repetition_penalty=1.25, no_repeat_ngram_size=2), dict(max_new_tokens=40, without_timestamps=True, repetition_penalty=1.3, no_repeat_ngram_size=3)])
def test_tiny_en_greedy_transcribe_token_exact(tiny_en, extra):
    gpu, cpu, mp, _ = tiny_en
    audio = synthetic_audio(123, 30.0)
    got_segs, got_info = gpu.transcribe(audio.copy(), **KW, **extra)
    got = list(got_segs)
    want_segs, want_info = _with_shim(mp, lambda: (lambda r: (list(r[0]), r[1]))(cpu.transcribe(audio.copy(), **KW, **extra)))
    assert len(want_segs) > 0 and sum(len(s.tokens) for s in want_segs) >= 20
    assert _tokens(got) == _tokens(want_segs)
    for g, w in zip(got, want_segs):
        assert (g.seek, round(g.start, 2), round(g.end, 2)) == (w.seek, round(w.start, 2), round(w.end, 2))
        assert abs(g.avg_logprob - w.avg_logprob) < 5e-3 and abs(g.no_speech_prob - w.no_speech_prob) < 5e-3
    assert got_info.language == want_info.language == "en"


def test_tiny_en_batched_pipeline_token_exact(tiny_en):
    gpu, cpu, mp, _ = tiny_en
    audio = np.concatenate([synthetic_audio(130 + i, 30.0) for i in range(3)])
    clips = [{"start": 30.0 * i, "end": 30.0 * (i + 1)} for i in range(3)]
    kw = dict(beam_size=1, batch_size=3, vad_filter=False, clip_timestamps=clips, max_new_tokens=32, repetition_penalty=1.2, no_repeat_ngram_size=3)
    got = list(T.BatchedInferencePipeline(gpu).transcribe(audio.copy(), **kw)[0])
    want = _with_shim(mp, lambda: list(T.BatchedInferencePipeline(cpu).transcribe(audio.copy(), **kw)[0]))
    assert len(want) >= 3
    assert _tokens(got) == _tokens(want)


def test_tiny_en_batched_clip_longer_than_30s(tiny_en):
    """A user clip of 41 s next to a 19 s one: the long clip's block goes log-mel (whole clip, CUDA kernel) -> features trimmed -> encode,
    the rest through the fused audio path; tokens as over the oracle shim."""
    gpu, cpu, mp, _ = tiny_en
    audio = np.concatenate([synthetic_audio(140, 30.0), 3.0 * synthetic_audio(141, 11.0), synthetic_audio(142, 19.0)])
    clips = [{"start": 0.0, "end": 41.0}, {"start": 41.0, "end": 60.0}]
    kw = dict(beam_size=1, batch_size=1, vad_filter=False, clip_timestamps=clips, max_new_tokens=24, repetition_penalty=1.2, no_repeat_ngram_size=3)
    got = list(T.BatchedInferencePipeline(gpu).transcribe(audio.copy(), **kw)[0])
    want = _with_shim(mp, lambda: list(T.BatchedInferencePipeline(cpu).transcribe(audio.copy(), **kw)[0]))
    assert len(want) >= 2
    assert _tokens(got) == _tokens(want)
