"""Small helpers with the reference's names (``faster_whisper/utils.py``).

``download_model`` needs the Hugging Face hub; there is no network in this environment, so it only resolves
local directories and otherwise raises with the hub repo the reference would have fetched (``utils.py:11-31``).
"""

from __future__ import annotations

import logging
import os
from typing import List, Optional, Union

from .config import MODEL_DIMS

_HUB_REPOS = {
    name: repo for name, repo in [
        ("tiny.en", "Systran/faster-whisper-tiny.en"), ("tiny", "Systran/faster-whisper-tiny"),
        ("base.en", "Systran/faster-whisper-base.en"), ("base", "Systran/faster-whisper-base"),
        ("small.en", "Systran/faster-whisper-small.en"), ("small", "Systran/faster-whisper-small"),
        ("medium.en", "Systran/faster-whisper-medium.en"), ("medium", "Systran/faster-whisper-medium"),
        ("large-v1", "Systran/faster-whisper-large-v1"), ("large-v2", "Systran/faster-whisper-large-v2"),
        ("large-v3", "Systran/faster-whisper-large-v3"), ("large", "Systran/faster-whisper-large-v3"),
        ("distil-large-v2", "Systran/faster-distil-whisper-large-v2"),
        ("distil-medium.en", "Systran/faster-distil-whisper-medium.en"),
        ("distil-small.en", "Systran/faster-distil-whisper-small.en"),
        ("distil-large-v3", "Systran/faster-distil-whisper-large-v3"),
        ("distil-large-v3.5", "distil-whisper/distil-large-v3.5-ct2"),
        ("large-v3-turbo", "mobiuslabsgmbh/faster-whisper-large-v3-turbo"),
        ("turbo", "mobiuslabsgmbh/faster-whisper-large-v3-turbo"),
    ]
}
assert set(_HUB_REPOS) == set(MODEL_DIMS)


def available_models() -> List[str]:
    """Names accepted as ``model_size_or_path``."""
    return list(_HUB_REPOS)


def get_logger() -> logging.Logger:
    return logging.getLogger("faster_whisper")


def download_model(size_or_id: str, output_dir: Optional[str] = None, local_files_only: bool = False,
                   cache_dir: Optional[str] = None, revision: Optional[str] = None,
                   use_auth_token: Optional[Union[str, bool]] = None) -> str:
    if os.path.isdir(size_or_id):
        return size_or_id
    repo = _HUB_REPOS.get(size_or_id, size_or_id if "/" in size_or_id else None)
    if repo is None:
        raise ValueError("Invalid model size '%s', expected one of: %s" % (size_or_id, ", ".join(_HUB_REPOS)))
    for root in (output_dir, cache_dir):
        if root and os.path.isdir(os.path.join(root, size_or_id)):
            return os.path.join(root, size_or_id)
    raise RuntimeError(
        f"cannot download '{repo}': no network access in this environment. Pass a local model directory, or build a "
        f"synthetic checkpoint with WhisperModel('{size_or_id}', synthetic_seed=0).")


def format_timestamp(seconds: float, always_include_hours: bool = False, decimal_marker: str = ".") -> str:
    assert seconds >= 0, "non-negative timestamp expected"
    total_ms = round(seconds * 1000.0)
    hours, rem = divmod(total_ms, 3_600_000)
    minutes, rem = divmod(rem, 60_000)
    secs, ms = divmod(rem, 1_000)
    head = f"{hours:02d}:" if always_include_hours or hours > 0 else ""
    return f"{head}{minutes:02d}:{secs:02d}{decimal_marker}{ms:03d}"


def get_end(segments: List[dict]) -> Optional[float]:
    """End time of the last word of the last segment that has words, else the last segment's end."""
    for seg in reversed(segments):
        for word in reversed(seg.get("words") or []):
            return word["end"]
    return segments[-1]["end"] if segments else None
