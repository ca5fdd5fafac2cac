"""Audio helpers with the reference's names (``faster_whisper/audio.py``).

``decode_audio`` is off the measured path (file decode is PyAV/FFmpeg, absent here);
it supports WAV through the standard library and defers to PyAV when that is installed.  ``pad_or_trim``
mirrors ``audio.py:111-123``.
"""

from __future__ import annotations

import wave
from typing import BinaryIO, Tuple, Union

import numpy as np


def _resample_linear(x: np.ndarray, src: int, dst: int) -> np.ndarray:
    if src == dst or x.size == 0:
        return x
    n = int(round(x.shape[0] * dst / src))
    pos = np.arange(n, dtype=np.float64) * (src / dst)
    return np.interp(pos, np.arange(x.shape[0]), x).astype(np.float32)


def decode_audio(input_file: Union[str, BinaryIO], sampling_rate: int = 16000, split_stereo: bool = False):
    """Returns float32 mono audio at `sampling_rate` (or a (left, right) tuple with split_stereo)."""
    try:
        with wave.open(input_file, "rb") as w:
            ch, width, rate, n = w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()
            raw = w.readframes(n)
    except (wave.Error, EOFError) as e:
        try:
            import av  # noqa: F401
        except ImportError:
            raise RuntimeError(
                "decode_audio: only PCM WAV input is supported without PyAV (FFmpeg) installed; "
                "pass a 16 kHz float32 NumPy array instead") from e
        return _decode_with_av(input_file, sampling_rate, split_stereo)
    if width == 2:
        pcm = np.frombuffer(raw, dtype="<i2").astype(np.float32) / 32768.0
    elif width == 4:
        pcm = np.frombuffer(raw, dtype="<i4").astype(np.float32) / 2147483648.0
    elif width == 1:
        pcm = (np.frombuffer(raw, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
    else:
        raise RuntimeError(f"unsupported WAV sample width {width}")
    pcm = pcm.reshape(-1, ch)
    if split_stereo:
        if ch != 2:
            raise RuntimeError("split_stereo needs a 2-channel input")
        return tuple(_resample_linear(pcm[:, i], rate, sampling_rate) for i in range(2))
    return _resample_linear(pcm.mean(axis=1).astype(np.float32), rate, sampling_rate)


def _decode_with_av(input_file, sampling_rate, split_stereo):  # pragma: no cover - PyAV is not in this image
    import av

    resampler = av.audio.resampler.AudioResampler(format="s16", layout="stereo" if split_stereo else "mono", rate=sampling_rate)
    parts = []
    with av.open(input_file, mode="r", metadata_errors="ignore") as container:
        for frame in container.decode(audio=0):
            for out in resampler.resample(frame):
                parts.append(out.to_ndarray().reshape(-1))
        for out in resampler.resample(None):
            parts.append(out.to_ndarray().reshape(-1))
    audio = np.concatenate(parts).astype(np.float32) / 32768.0 if parts else np.zeros(0, np.float32)
    if split_stereo:
        return audio[0::2], audio[1::2]
    return audio


def pad_or_trim(array: np.ndarray, length: int = 3000, *, axis: int = -1) -> np.ndarray:
    """Zero-pad or cut `axis` to `length` frames (3000 = 30 s, what the encoder expects)."""
    n = array.shape[axis]
    if n > length:
        index = [slice(None)] * array.ndim
        index[axis] = slice(0, length)
        return array[tuple(index)]
    if n < length:
        widths = [(0, 0)] * array.ndim
        widths[axis] = (0, length - n)
        return np.pad(array, widths)
    return array
