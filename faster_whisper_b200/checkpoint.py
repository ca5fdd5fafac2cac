"""Model directories for this engine.

The reference loads a CTranslate2 directory (``model.bin`` + ``config.json`` + ``tokenizer.json`` +
``preprocessor_config.json``; ``faster_whisper/utils.py:91-97``, ``transcribe.py:689-710``).  No such checkpoint
exists offline, so the native on-disk format here is the plainest possible: ``weights.npz`` (OpenAI-Whisper
state-dict names, float16/float32) + ``b2w_config.json`` (geometry) and, optionally, ``tokenizer.json`` /
``preprocessor_config.json`` exactly as in a CTranslate2 directory.  Directories holding a CTranslate2 ``model.bin`` are
read through ``ct2_format.py`` (layout restated, unpinned); directories holding a ``transformers``
checkpoint (``config.json`` + ``model.safetensors``) through ``hf_format.py`` (pinned against transformers' own writer).
"""

from __future__ import annotations

import io
import json
import os
from typing import Dict, Optional, Tuple

import numpy as np

from .config import MODEL_DIMS, WhisperDims


def save_model_dir(path: str, dims: WhisperDims, weights: Dict[str, np.ndarray], tokenizer=None,
                   dtype=np.float16) -> None:
    os.makedirs(path, exist_ok=True)
    np.savez(os.path.join(path, "weights.npz"), **{k: v.astype(dtype) for k, v in weights.items()})
    with open(os.path.join(path, "b2w_config.json"), "w", encoding="utf-8") as f:
        json.dump(dims.to_dict(), f)
    with open(os.path.join(path, "preprocessor_config.json"), "w", encoding="utf-8") as f:
        json.dump({"feature_size": dims.n_mels, "sampling_rate": 16000, "hop_length": 160, "chunk_length": 30,
                   "n_fft": 400}, f)
    if tokenizer is not None:
        tokenizer.save(os.path.join(path, "tokenizer.json"))


def load_model_dir(path: str, files: Optional[dict] = None) -> Tuple[WhisperDims, Dict[str, np.ndarray]]:
    def read(name):
        if files and name in files:
            blob = files[name]
            return blob.read() if hasattr(blob, "read") else blob
        p = os.path.join(path, name)
        if not os.path.isfile(p):
            return None
        with open(p, "rb") as f:
            return f.read()

    def exists(name):  # never consumes a file-like entry and never reads a multi-GB blob just to probe
        return bool(files and name in files) or bool(path and os.path.isfile(os.path.join(path, name)))

    if not (exists("b2w_config.json") and exists("weights.npz")):
        if exists("model.bin"):
            # a CTranslate2 directory, as the reference loads it (ct2_format.py; layout restated, unpinned)
            from .ct2_format import load_ct2_dir

            dims, weights, _ = load_ct2_dir(path, files)
            return dims, weights
        from .hf_format import is_hf_dir, load_hf_dir

        if is_hf_dir(path, files):
            # a transformers checkpoint (config.json + model.safetensors), pinned by tests/test_hf_loader.py
            dims, weights, _ = load_hf_dir(path, files)
            return dims, weights
        raise RuntimeError(f"Unable to open file 'weights.npz' in model '{path}'")
    cfg = read("b2w_config.json")
    wts = read("weights.npz")
    dims = WhisperDims(**json.loads(cfg))
    with np.load(io.BytesIO(wts)) as z:
        weights = {k: z[k] for k in z.files}
    return dims, weights


def read_alignment_heads(path: str, files: Optional[dict] = None):
    """``alignment_heads`` of a model directory's ``config.json`` ([[layer, head], ...]) or None."""
    blob = None
    if files and "config.json" in files:
        blob = files["config.json"]
        blob = blob.read() if hasattr(blob, "read") else blob
    elif path and os.path.isfile(os.path.join(path, "config.json")):
        with open(os.path.join(path, "config.json"), "rb") as f:
            blob = f.read()
    if not blob:
        return None
    try:
        heads = json.loads(blob).get("alignment_heads")
    except ValueError:
        return None
    return [tuple(int(x) for x in p) for p in heads] if heads else None
