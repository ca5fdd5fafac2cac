"""Model directories for this engine.

The reference loads a CTranslate2 directory (``model.bin`` + ``config.json`` + ``tokenizer.json`` +
``preprocessor_config.json``; ``faster_whisper/utils.py:91-97``, ``transcribe.py:689-710``).  No such checkpoint
exists offline, so the native on-disk format here is the plainest possible: ``weights.npz`` (OpenAI-Whisper
state-dict names, float16/float32) + ``b2w_config.json`` (geometry) and, optionally, ``tokenizer.json`` /
``preprocessor_config.json`` exactly as in a CTranslate2 directory.  A ``model.bin`` reader is SURVEY.md §8(f)
row 1 ("next").
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

    cfg = read("b2w_config.json")
    wts = read("weights.npz")
    if cfg is None or wts is None:
        if read("model.bin") is not None:
            raise RuntimeError(f"{path} holds a CTranslate2 model.bin; the model.bin reader is not implemented yet "
                               "(convert to weights.npz + b2w_config.json with faster_whisper_b200.checkpoint.save_model_dir)")
        raise RuntimeError(f"Unable to open file 'weights.npz' in model '{path}'")
    dims = WhisperDims(**json.loads(cfg))
    with np.load(io.BytesIO(wts)) as z:
        weights = {k: z[k] for k in z.files}
    return dims, weights
