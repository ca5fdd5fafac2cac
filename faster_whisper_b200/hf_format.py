"""Reader for a Hugging Face ``transformers`` Whisper checkpoint directory (second half).

``config.json`` (``WhisperConfig``: d_model, encoder_layers, decoder_layers, *_attention_heads, num_mel_bins, vocab_size,
max_source_positions, max_target_positions) + ``model.safetensors`` (or sharded ``model-0000x-of-0000y.safetensors`` with
``model.safetensors.index.json``).  The safetensors container is parsed here directly (u64 little-endian header length, JSON
header ``{name: {dtype, shape, data_offsets}}``, raw little-endian data) so that the product needs neither torch nor the
safetensors package; tensors come back under the OpenAI-Whisper state-dict names the engine consumes
(``decoder.blocks.0.attn.query.weight`` ...), which is also what the CTranslate2 converter starts from.

PINNED: tests/test_hf_loader.py loads a directory written by ``transformers.WhisperForConditionalGeneration.save_pretrained``
(an independent writer) and recovers every tensor bit for bit.
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, Optional, Tuple

import numpy as np

from .config import WhisperDims

_ST_DTYPES = {"F32": np.float32, "F16": np.float16, "F64": np.float64, "I64": np.int64, "I32": np.int32, "I8": np.int8, "U8": np.uint8}


def _bf16_to_f32(raw: np.ndarray) -> np.ndarray:
    return (raw.astype(np.uint32) << 16).view(np.float32)


def read_safetensors(blob) -> Dict[str, np.ndarray]:
    """`blob`: path, bytes or a binary file object -> {name: array} (bfloat16 widened to float32)."""
    if isinstance(blob, (bytes, bytearray, memoryview)):
        data = bytes(blob)
    elif hasattr(blob, "read"):
        data = blob.read()
    else:
        with open(blob, "rb") as f:
            data = f.read()
    if len(data) < 8:
        raise ValueError("not a safetensors file (too short)")
    (hlen,) = struct.unpack("<Q", data[:8])
    if hlen <= 0 or 8 + hlen > len(data):
        raise ValueError("not a safetensors file (bad header length)")
    header = json.loads(data[8 : 8 + hlen].decode("utf-8"))
    base = 8 + hlen
    out: Dict[str, np.ndarray] = {}
    for name, meta in header.items():
        if name == "__metadata__":
            continue
        a, b = meta["data_offsets"]
        shape = tuple(int(x) for x in meta["shape"])
        dt = meta["dtype"]
        if dt == "BF16":
            arr = _bf16_to_f32(np.frombuffer(data, dtype=np.uint16, count=(b - a) // 2, offset=base + a))
        elif dt in _ST_DTYPES:
            npdt = np.dtype(_ST_DTYPES[dt])
            arr = np.frombuffer(data, dtype=npdt, count=(b - a) // npdt.itemsize, offset=base + a)
        else:
            raise ValueError(f"safetensors dtype {dt} of '{name}' is not supported")
        out[name] = arr.reshape(shape)
    return out


_ATTN = {"q_proj": "query", "k_proj": "key", "v_proj": "value", "out_proj": "out"}


def hf_to_openai_name(name: str) -> Optional[str]:
    """``model.decoder.layers.3.encoder_attn.q_proj.weight`` -> ``decoder.blocks.3.cross_attn.query.weight``; None = not needed."""
    if name.startswith("model."):
        name = name[len("model."):]
    if name == "proj_out.weight":
        return None  # tied to decoder.embed_tokens
    parts = name.split(".")
    side = parts[0]
    if side not in ("encoder", "decoder"):
        return None
    rest = parts[1:]
    if rest[0] in ("conv1", "conv2") and side == "encoder":
        return f"encoder.{rest[0]}.{rest[1]}"
    if rest[0] == "embed_positions":
        return f"{side}.positional_embedding"
    if rest[0] == "embed_tokens":
        return "decoder.token_embedding.weight"
    if rest[0] == "layer_norm":
        return ("encoder.ln_post." if side == "encoder" else "decoder.ln.") + rest[1]
    if rest[0] == "layers":
        i, sub = rest[1], rest[2:]
        p = f"{side}.blocks.{i}"
        if sub[0] in ("self_attn", "encoder_attn"):
            a = "attn" if sub[0] == "self_attn" else "cross_attn"
            return f"{p}.{a}.{_ATTN[sub[1]]}.{sub[2]}"
        if sub[0] == "self_attn_layer_norm":
            return f"{p}.attn_ln.{sub[1]}"
        if sub[0] == "encoder_attn_layer_norm":
            return f"{p}.cross_attn_ln.{sub[1]}"
        if sub[0] == "final_layer_norm":
            return f"{p}.mlp_ln.{sub[1]}"
        if sub[0] == "fc1":
            return f"{p}.mlp.0.{sub[1]}"
        if sub[0] == "fc2":
            return f"{p}.mlp.2.{sub[1]}"
    return None


def is_hf_dir(path: str, files: Optional[dict] = None) -> bool:
    names = set(files or ())
    if path and os.path.isdir(path):
        names |= set(os.listdir(path))
    return "config.json" in names and ("model.safetensors" in names or "model.safetensors.index.json" in names)


def load_hf_dir(path: str, files: Optional[dict] = None) -> Tuple[WhisperDims, Dict[str, np.ndarray], dict]:
    """-> (dims, OpenAI-named float32/float16 weights, raw config dict)."""
    def read(name):
        if files and name in files:
            blob = files[name]
            if hasattr(blob, "read"):
                if hasattr(blob, "seek"):
                    blob.seek(0)
                return blob.read()
            return blob
        with open(os.path.join(path, name), "rb") as f:
            return f.read()

    cfg = json.loads(read("config.json"))
    if "d_model" not in cfg:
        raise ValueError("config.json is not a transformers WhisperConfig (no d_model)")
    shards = ["model.safetensors"]
    try:
        index = json.loads(read("model.safetensors.index.json"))
        shards = sorted(set(index["weight_map"].values()))
    except (FileNotFoundError, KeyError):
        pass
    weights: Dict[str, np.ndarray] = {}
    for shard in shards:
        for name, arr in read_safetensors(read(shard)).items():
            ours = hf_to_openai_name(name)
            if ours is not None:
                weights[ours] = np.ascontiguousarray(arr)
    d = int(cfg["d_model"])
    dims = WhisperDims(str(cfg.get("_name_or_path") or "hf-whisper"), int(cfg["num_mel_bins"]), d, int(cfg["encoder_attention_heads"]),
                       int(cfg["encoder_layers"]), d, int(cfg["decoder_attention_heads"]), int(cfg["decoder_layers"]), int(cfg["vocab_size"]),
                       int(cfg.get("max_source_positions", 1500)), int(cfg.get("max_target_positions", 448)))
    if "encoder.positional_embedding" not in weights:
        # transformers keeps the sinusoids as a (non-persistent in some versions) parameter: rebuild Whisper's fixed table
        from .synthetic import sinusoids

        weights["encoder.positional_embedding"] = sinusoids(dims.n_audio_ctx, d)
    need = ["decoder.token_embedding.weight", "decoder.positional_embedding", "encoder.conv1.weight", "decoder.ln.weight"]
    missing = [n for n in need if n not in weights]
    if missing:
        raise ValueError("incomplete Whisper checkpoint, missing: " + ", ".join(missing))
    return dims, weights, cfg
