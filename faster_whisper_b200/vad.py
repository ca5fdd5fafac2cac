"""Chunk bookkeeping around voice-activity detection (``faster_whisper/vad.py``).

The Silero VAD network itself (ONNX, CPU LSTM) is outside the accelerated path and
onnxruntime is not installed here, so ``get_speech_timestamps`` raises unless onnxruntime and the model asset
are available; the pure-Python pieces the batched pipeline needs — ``VadOptions``, ``collect_chunks``,
``SpeechTimestampsMap`` — behave like ``vad.py:14-42,186-285``.
"""

from __future__ import annotations

import bisect
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np


@dataclass
class VadOptions:
    threshold: float = 0.5
    neg_threshold: float = None
    min_speech_duration_ms: int = 0
    max_speech_duration_s: float = float("inf")
    min_silence_duration_ms: int = 2000
    speech_pad_ms: int = 400


def get_speech_timestamps(audio: np.ndarray, vad_options: Optional[VadOptions] = None, sampling_rate: int = 16000, **kwargs):
    raise RuntimeError(
        "VAD needs onnxruntime and the Silero model asset, neither of which ships with this engine; "
        "use vad_filter=False and pass clip_timestamps (the chunked 30 s path) instead")


def collect_chunks(audio: np.ndarray, chunks: List[dict], sampling_rate: int = 16000,
                   max_duration: float = float("inf")) -> Tuple[List[np.ndarray], List[Dict[str, float]]]:
    """Greedily packs speech spans (sample offsets) into pieces of at most `max_duration` seconds."""
    if not chunks:
        return [np.array([], dtype=np.float32)], [{"offset": 0, "duration": 0, "segments": []}]
    limit = max_duration * sampling_rate
    pieces, metas = [], []
    spans, samples, consumed = [], 0, 0
    buf: List[np.ndarray] = []

    def flush():
        nonlocal spans, samples, consumed, buf
        pieces.append(np.concatenate(buf) if buf else np.array([], dtype=np.float32))
        metas.append({"offset": consumed / sampling_rate, "duration": samples / sampling_rate, "segments": spans})
        consumed += samples

    for span in chunks:
        n = span["end"] - span["start"]
        if samples + n > limit:
            flush()
            # the reference starts the next piece with this span's audio but does not list the span itself
            spans, buf, samples = [], [audio[span["start"] : span["end"]]], n
        else:
            spans.append(span)
            buf.append(audio[span["start"] : span["end"]])
            samples += n
    flush()
    return pieces, metas


class SpeechTimestampsMap:
    """Maps times on the silence-removed axis back to the original recording."""

    def __init__(self, chunks: List[dict], sampling_rate: int, time_precision: int = 2):
        self.sampling_rate = sampling_rate
        self.time_precision = time_precision
        self.chunk_end_sample: List[int] = []
        self.total_silence_before: List[float] = []
        cursor, removed = 0, 0
        for c in chunks:
            removed += c["start"] - cursor
            cursor = c["end"]
            self.chunk_end_sample.append(c["end"] - removed)
            self.total_silence_before.append(removed / sampling_rate)

    def get_chunk_index(self, time: float, is_end: bool = False) -> int:
        sample = int(time * self.sampling_rate)
        if is_end and sample in self.chunk_end_sample:
            return self.chunk_end_sample.index(sample)
        return min(bisect.bisect(self.chunk_end_sample, sample), len(self.chunk_end_sample) - 1)

    def get_original_time(self, time: float, chunk_index: Optional[int] = None, is_end: bool = False) -> float:
        if chunk_index is None:
            chunk_index = self.get_chunk_index(time, is_end)
        return round(self.total_silence_before[chunk_index] + time, self.time_precision)
