"""Small helpers with the reference's names (``faster_whisper/utils.py``).

``download_model`` resolves local directories and otherwise goes to the Hugging Face hub like the reference
(``utils.py:49-116``); offline it raises with the hub repo it tried (``utils.py:11-31`` names them).
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
    """Resolves a model size / hub repo id to a local CTranslate2 model directory (reference ``utils.py:49-116``).

    A local directory is returned as is; otherwise the files the engine reads (``model.bin``, ``config.json``,
    ``preprocessor_config.json``, ``tokenizer.json``, ``vocabulary.*``) are fetched with
    ``huggingface_hub.snapshot_download`` exactly as the reference does — honouring ``output_dir``, ``cache_dir``,
    ``revision``, ``use_auth_token`` and ``local_files_only``.  Without the hub package, or when the hub cannot be reached
    and nothing is cached, a RuntimeError names the repository that would have been fetched.
    """
    if os.path.isdir(size_or_id):
        return size_or_id
    repo = size_or_id if "/" in size_or_id else _HUB_REPOS.get(size_or_id)
    if repo is None:
        raise ValueError("Invalid model size '%s', expected one of: %s" % (size_or_id, ", ".join(_HUB_REPOS)))
    for root in (output_dir, cache_dir):  # a plain directory laid out by hand next to the caches
        if root and os.path.isfile(os.path.join(root, size_or_id, "model.bin")):
            return os.path.join(root, size_or_id)
    try:
        import huggingface_hub
    except ImportError as e:  # pragma: no cover - the image ships it
        raise RuntimeError(f"cannot fetch '{repo}': huggingface_hub is not installed; pass a local model directory") from e
    options = dict(local_files_only=local_files_only, revision=revision,
                   allow_patterns=["config.json", "preprocessor_config.json", "model.bin", "tokenizer.json", "vocabulary.*"])
    if output_dir is not None:
        options["local_dir"] = output_dir
    if cache_dir is not None:
        options["cache_dir"] = cache_dir
    if use_auth_token is not None:
        options["token"] = use_auth_token
    try:
        from tqdm.auto import tqdm as _tqdm

        class _QuietTqdm(_tqdm):
            def __init__(self, *a, **k):
                k["disable"] = True
                super().__init__(*a, **k)

        options["tqdm_class"] = _QuietTqdm
    except Exception:  # noqa: BLE001 - progress bars are cosmetic
        pass
    try:
        return huggingface_hub.snapshot_download(repo, **options)
    except Exception as e:  # noqa: BLE001 - offline / not cached / auth: say what was attempted
        raise RuntimeError(
            f"cannot fetch '{repo}' from the Hugging Face hub ({type(e).__name__}: {e}). Pass a local model directory, or "
            f"build a synthetic checkpoint with WhisperModel('{size_or_id}', synthetic_seed=0).") from e


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
