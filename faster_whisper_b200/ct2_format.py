"""Reader (and writer) for the CTranslate2 model directory the reference loads.

``WhisperModel(model_path)`` in the reference hands the directory to ``ctranslate2.models.Whisper`` (``transcribe.py:689-698``)
and itself reads ``tokenizer.json`` / ``preprocessor_config.json`` next to it (``:700-745``, ``utils.py:91-97``).  The
directory holds

    model.bin       u32 binary version; [v>=2] string spec name, u32 spec revision; u32 variable count; per variable:
                    string name, u8 rank, u32 dims[rank], [v>=4] u8 dtype id + u32 byte count (else u8 item size + u32 count),
                    raw little-endian data; [v>=3] u32 alias count + (string alias, string target) pairs.
                    A string is a u16 length (terminating NUL included) followed by the bytes.
    config.json     alignment_heads [[layer, head]...], lang_ids, suppress_ids, suppress_ids_begin, ...

No such file exists in this build environment (no network), so the layout above and the variable names below are restated
from CTranslate2 4.x's published converter/spec conventions and are UNPINNED; `write_model_bin` produces the same layout so
that synthetic checkpoints round-trip (tests/test_host_logic.py) and a real ``Systran/faster-whisper-*`` directory can be
tried as soon as one is available.  Variable naming (WhisperSpec): nested attributes joined by ``/``, list items suffixed
``_i``; self-attention ``linear_0`` = fused QKV, ``linear_1`` = output; cross-attention (``attention``) ``linear_0`` = Q,
``linear_1`` = fused KV, ``linear_2`` = output; int8/int16 weights carry a per-output-row ``weight_scale`` with
w = q / scale.
"""
from __future__ import annotations

import io
import json
import os
import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

from .config import WhisperDims

DTYPES = {0: np.float32, 1: np.int8, 2: np.int16, 3: np.int32, 4: np.float16}
DTYPE_IDS = {np.dtype(v): k for k, v in DTYPES.items()}
BFLOAT16_ID = 5
BINARY_VERSION = 6


def _read_string(f) -> str:
    (n,) = struct.unpack("<H", f.read(2))
    raw = f.read(n)
    return raw.rstrip(b"\0").decode("utf-8")


def _write_string(f, s: str) -> None:
    raw = s.encode("utf-8") + b"\0"
    f.write(struct.pack("<H", len(raw)))
    f.write(raw)


def read_model_bin(blob) -> Tuple[str, int, Dict[str, np.ndarray], Dict[str, str]]:
    """-> (spec name, spec revision, variables, aliases).  `blob`: path, bytes or a binary file object."""
    if isinstance(blob, (bytes, bytearray)):
        f = io.BytesIO(blob)
    elif hasattr(blob, "read"):
        f = blob
    else:
        f = open(blob, "rb")
    try:
        (version,) = struct.unpack("<I", f.read(4))
        if not 1 <= version <= 16:
            raise ValueError(f"not a CTranslate2 model.bin (binary version {version})")
        spec, revision = "", 1
        if version >= 2:
            spec = _read_string(f)
            (revision,) = struct.unpack("<I", f.read(4))
        (count,) = struct.unpack("<I", f.read(4))
        variables: Dict[str, np.ndarray] = {}
        for _ in range(count):
            name = _read_string(f)
            (rank,) = struct.unpack("<B", f.read(1))
            dims = struct.unpack(f"<{rank}I", f.read(4 * rank)) if rank else ()
            if version >= 4:
                dtype_id, nbytes = struct.unpack("<BI", f.read(5))
            else:
                item, n = struct.unpack("<BI", f.read(5))
                dtype_id, nbytes = {4: 0, 2: 2, 1: 1}[item], item * n
            raw = f.read(nbytes)
            if len(raw) != nbytes:
                raise ValueError(f"model.bin is truncated inside variable {name!r}")
            if dtype_id == BFLOAT16_ID:
                arr = (np.frombuffer(raw, dtype="<u2").astype(np.uint32) << 16).view(np.float32)
            elif dtype_id in DTYPES:
                arr = np.frombuffer(raw, dtype=np.dtype(DTYPES[dtype_id]).newbyteorder("<"))
            else:
                raise ValueError(f"variable {name!r} has unknown dtype id {dtype_id}")
            variables[name] = arr.reshape(dims)
        aliases: Dict[str, str] = {}
        if version >= 3:
            head = f.read(4)
            if len(head) == 4:
                for _ in range(struct.unpack("<I", head)[0]):
                    alias = _read_string(f)
                    aliases[alias] = _read_string(f)
        return spec, revision, variables, aliases
    finally:
        if not isinstance(blob, (bytes, bytearray)) and not hasattr(blob, "read"):
            f.close()


def write_model_bin(path_or_file, variables: Dict[str, np.ndarray], aliases: Optional[Dict[str, str]] = None,
                    spec: str = "WhisperSpec", revision: int = 3) -> None:
    f = path_or_file if hasattr(path_or_file, "write") else open(path_or_file, "wb")
    try:
        f.write(struct.pack("<I", BINARY_VERSION))
        _write_string(f, spec)
        f.write(struct.pack("<I", revision))
        f.write(struct.pack("<I", len(variables)))
        for name in sorted(variables):
            arr = np.ascontiguousarray(variables[name])
            _write_string(f, name)
            f.write(struct.pack("<B", arr.ndim))
            if arr.ndim:
                f.write(struct.pack(f"<{arr.ndim}I", *arr.shape))
            f.write(struct.pack("<BI", DTYPE_IDS[arr.dtype], arr.nbytes))
            f.write(arr.tobytes())
        aliases = aliases or {}
        f.write(struct.pack("<I", len(aliases)))
        for alias in sorted(aliases):
            _write_string(f, alias)
            _write_string(f, aliases[alias])
    finally:
        if not hasattr(path_or_file, "write"):
            f.close()


# ---- CTranslate2 WhisperSpec variables <-> OpenAI-style state dict (the names csrc/engine.cu:build_model reads) ------------
def _dense(variables: Dict[str, np.ndarray], aliases: Dict[str, str], prefix: str) -> np.ndarray:
    """Weight of a LinearSpec/Conv1DSpec/EmbeddingsSpec as float32, de-quantised when it is stored as int8/int16."""
    name = prefix + "/weight"
    name = aliases.get(name, name)
    w = variables[name]
    if w.dtype in (np.int8, np.int16):
        scale = variables[name[: -len("weight")] + "weight_scale"].astype(np.float32)
        return w.astype(np.float32) / scale.reshape((-1,) + (1,) * (w.ndim - 1))
    return w.astype(np.float32)


def _vec(variables, aliases, name) -> np.ndarray:
    return variables[aliases.get(name, name)].astype(np.float32)


def ct2_to_state_dict(variables: Dict[str, np.ndarray], aliases: Optional[Dict[str, str]] = None, name: str = "ct2"):
    """-> (WhisperDims, {OpenAI-style name: float32 array})."""
    aliases = aliases or {}
    v = variables

    def has(n):
        return aliases.get(n, n) in v

    def count(fmt):
        i = 0
        while has(fmt.format(i) + "/self_attention/linear_0/weight"):
            i += 1
        return i

    n_enc, n_dec = count("encoder/layer_{}"), count("decoder/layer_{}")
    if n_enc == 0 or n_dec == 0:
        raise ValueError("model.bin does not look like a WhisperSpec (no encoder/decoder layers); variables: "
                         + ", ".join(sorted(v)[:8]) + " ...")
    conv1 = _dense(v, aliases, "encoder/conv1")
    d, n_mels = conv1.shape[0], conv1.shape[1]
    emb = _dense(v, aliases, "decoder/embeddings")
    dt = emb.shape[1]
    heads_e = int(np.asarray(v["encoder/num_heads"]).reshape(-1)[0]) if "encoder/num_heads" in v else d // 64
    heads_d = int(np.asarray(v["decoder/num_heads"]).reshape(-1)[0]) if "decoder/num_heads" in v else dt // 64
    enc_pos = _vec(v, aliases, "encoder/position_encodings/encodings")
    dec_pos = _vec(v, aliases, "decoder/position_encodings/encodings")
    dims = WhisperDims(name=name, n_mels=int(n_mels), n_audio_ctx=int(enc_pos.shape[0]), n_audio_state=int(d), n_audio_head=heads_e,
                       n_audio_layer=n_enc, n_vocab=int(emb.shape[0]), n_text_ctx=int(dec_pos.shape[0]), n_text_state=int(dt),
                       n_text_head=heads_d, n_text_layer=n_dec)
    w: Dict[str, np.ndarray] = {}
    for i in (1, 2):
        w[f"encoder.conv{i}.weight"] = _dense(v, aliases, f"encoder/conv{i}")
        w[f"encoder.conv{i}.bias"] = _vec(v, aliases, f"encoder/conv{i}/bias")
    w["encoder.positional_embedding"] = enc_pos
    w["encoder.ln_post.weight"] = _vec(v, aliases, "encoder/layer_norm/gamma")
    w["encoder.ln_post.bias"] = _vec(v, aliases, "encoder/layer_norm/beta")
    w["decoder.token_embedding.weight"] = emb
    w["decoder.positional_embedding"] = dec_pos
    w["decoder.ln.weight"] = _vec(v, aliases, "decoder/layer_norm/gamma")
    w["decoder.ln.bias"] = _vec(v, aliases, "decoder/layer_norm/beta")

    def self_attention(src, dst, width):
        w[dst + "attn_ln.weight"] = _vec(v, aliases, src + "/layer_norm/gamma")
        w[dst + "attn_ln.bias"] = _vec(v, aliases, src + "/layer_norm/beta")
        qkv = _dense(v, aliases, src + "/linear_0")
        b = _vec(v, aliases, src + "/linear_0/bias")
        for k, nm in enumerate(("query", "key", "value")):
            w[f"{dst}attn.{nm}.weight"] = qkv[k * width : (k + 1) * width]
            if nm != "key":
                w[f"{dst}attn.{nm}.bias"] = b[k * width : (k + 1) * width]
        w[dst + "attn.out.weight"] = _dense(v, aliases, src + "/linear_1")
        w[dst + "attn.out.bias"] = _vec(v, aliases, src + "/linear_1/bias")

    def ffn(src, dst):
        w[dst + "mlp_ln.weight"] = _vec(v, aliases, src + "/layer_norm/gamma")
        w[dst + "mlp_ln.bias"] = _vec(v, aliases, src + "/layer_norm/beta")
        w[dst + "mlp.0.weight"] = _dense(v, aliases, src + "/linear_0")
        w[dst + "mlp.0.bias"] = _vec(v, aliases, src + "/linear_0/bias")
        w[dst + "mlp.2.weight"] = _dense(v, aliases, src + "/linear_1")
        w[dst + "mlp.2.bias"] = _vec(v, aliases, src + "/linear_1/bias")

    for i in range(n_enc):
        self_attention(f"encoder/layer_{i}/self_attention", f"encoder.blocks.{i}.", d)
        ffn(f"encoder/layer_{i}/ffn", f"encoder.blocks.{i}.")
    for i in range(n_dec):
        dst = f"decoder.blocks.{i}."
        self_attention(f"decoder/layer_{i}/self_attention", dst, dt)
        src = f"decoder/layer_{i}/attention"
        w[dst + "cross_attn_ln.weight"] = _vec(v, aliases, src + "/layer_norm/gamma")
        w[dst + "cross_attn_ln.bias"] = _vec(v, aliases, src + "/layer_norm/beta")
        w[dst + "cross_attn.query.weight"] = _dense(v, aliases, src + "/linear_0")
        w[dst + "cross_attn.query.bias"] = _vec(v, aliases, src + "/linear_0/bias")
        kv = _dense(v, aliases, src + "/linear_1")
        kvb = _vec(v, aliases, src + "/linear_1/bias")
        w[dst + "cross_attn.key.weight"] = kv[:dt]
        w[dst + "cross_attn.value.weight"] = kv[dt:]
        w[dst + "cross_attn.value.bias"] = kvb[dt:]
        w[dst + "cross_attn.out.weight"] = _dense(v, aliases, src + "/linear_2")
        w[dst + "cross_attn.out.bias"] = _vec(v, aliases, src + "/linear_2/bias")
        ffn(f"decoder/layer_{i}/ffn", dst)
    return dims, w


def state_dict_to_ct2(dims: WhisperDims, weights: Dict[str, np.ndarray], dtype=np.float16, quantize_int8: bool = False):
    """The inverse (for exporting synthetic checkpoints and for the round-trip test).  -> (variables, aliases)."""
    v: Dict[str, np.ndarray] = {}

    def put_dense(prefix, w, bias=None):
        w = np.asarray(w, np.float32)
        if quantize_int8 and w.ndim == 2:
            scale = 127.0 / np.maximum(np.abs(w).max(axis=1), 1e-12)
            v[prefix + "/weight"] = np.clip(np.rint(w * scale[:, None]), -127, 127).astype(np.int8)
            v[prefix + "/weight_scale"] = scale.astype(np.float32)
        else:
            v[prefix + "/weight"] = w.astype(dtype)
        if bias is not None:
            v[prefix + "/bias"] = np.asarray(bias, np.float32).astype(dtype)

    def put_ln(prefix, g, b):
        v[prefix + "/gamma"] = np.asarray(g, np.float32).astype(dtype)
        v[prefix + "/beta"] = np.asarray(b, np.float32).astype(dtype)

    W = weights
    for i in (1, 2):
        put_dense(f"encoder/conv{i}", W[f"encoder.conv{i}.weight"], W[f"encoder.conv{i}.bias"])
    v["encoder/position_encodings/encodings"] = np.asarray(W["encoder.positional_embedding"], np.float32).astype(dtype)
    v["encoder/num_heads"] = np.asarray(dims.n_audio_head, np.int16)
    v["decoder/num_heads"] = np.asarray(dims.n_text_head, np.int16)
    put_ln("encoder/layer_norm", W["encoder.ln_post.weight"], W["encoder.ln_post.bias"])
    put_dense("decoder/embeddings", W["decoder.token_embedding.weight"])
    v["decoder/position_encodings/encodings"] = np.asarray(W["decoder.positional_embedding"], np.float32).astype(dtype)
    put_ln("decoder/layer_norm", W["decoder.ln.weight"], W["decoder.ln.bias"])

    def self_attention(dst, src, width):
        put_ln(dst + "/layer_norm", W[src + "attn_ln.weight"], W[src + "attn_ln.bias"])
        qkv = np.concatenate([W[f"{src}attn.{n}.weight"] for n in ("query", "key", "value")])
        b = np.concatenate([W[src + "attn.query.bias"], np.zeros(width, np.float32), W[src + "attn.value.bias"]])
        put_dense(dst + "/linear_0", qkv, b)
        put_dense(dst + "/linear_1", W[src + "attn.out.weight"], W[src + "attn.out.bias"])

    def ffn(dst, src):
        put_ln(dst + "/layer_norm", W[src + "mlp_ln.weight"], W[src + "mlp_ln.bias"])
        put_dense(dst + "/linear_0", W[src + "mlp.0.weight"], W[src + "mlp.0.bias"])
        put_dense(dst + "/linear_1", W[src + "mlp.2.weight"], W[src + "mlp.2.bias"])

    for i in range(dims.n_audio_layer):
        self_attention(f"encoder/layer_{i}/self_attention", f"encoder.blocks.{i}.", dims.n_audio_state)
        ffn(f"encoder/layer_{i}/ffn", f"encoder.blocks.{i}.")
    dt = dims.n_text_state
    for i in range(dims.n_text_layer):
        src = f"decoder.blocks.{i}."
        self_attention(f"decoder/layer_{i}/self_attention", src, dt)
        dst = f"decoder/layer_{i}/attention"
        put_ln(dst + "/layer_norm", W[src + "cross_attn_ln.weight"], W[src + "cross_attn_ln.bias"])
        put_dense(dst + "/linear_0", W[src + "cross_attn.query.weight"], W[src + "cross_attn.query.bias"])
        put_dense(dst + "/linear_1", np.concatenate([W[src + "cross_attn.key.weight"], W[src + "cross_attn.value.weight"]]),
                  np.concatenate([np.zeros(dt, np.float32), W[src + "cross_attn.value.bias"]]))
        put_dense(dst + "/linear_2", W[src + "cross_attn.out.weight"], W[src + "cross_attn.out.bias"])
        ffn(f"decoder/layer_{i}/ffn", src)
    aliases = {"decoder/projection/weight": "decoder/embeddings/weight"}
    if quantize_int8:
        aliases["decoder/projection/weight_scale"] = "decoder/embeddings/weight_scale"
    return v, aliases


def load_ct2_dir(path: str, files: Optional[dict] = None):
    """-> (WhisperDims, weights, config dict) from a CTranslate2 model directory (or the in-memory `files` mapping the
    reference also accepts, transcribe.py:689-698)."""

    def read(name):
        if files and name in files:
            blob = files[name]
            return blob.read() if hasattr(blob, "read") else blob
        p = os.path.join(path, name)
        if not os.path.isfile(p):
            return None
        with open(p, "rb") as f:
            return f.read()

    blob = read("model.bin")
    if blob is None:
        raise RuntimeError(f"Unable to open file 'model.bin' in model '{path}'")
    _, _, variables, aliases = read_model_bin(blob)
    dims, weights = ct2_to_state_dict(variables, aliases, name=os.path.basename(os.path.normpath(path)) or "ct2")
    cfg_raw = read("config.json")
    config = json.loads(cfg_raw) if cfg_raw else {}
    return dims, weights, config


def save_ct2_dir(path: str, dims: WhisperDims, weights: Dict[str, np.ndarray], alignment_heads: Optional[List[Tuple[int, int]]] = None,
                 tokenizer=None, dtype=np.float16, quantize_int8: bool = False) -> None:
    os.makedirs(path, exist_ok=True)
    variables, aliases = state_dict_to_ct2(dims, weights, dtype=dtype, quantize_int8=quantize_int8)
    write_model_bin(os.path.join(path, "model.bin"), variables, aliases)
    cfg = {"alignment_heads": [list(p) for p in (alignment_heads or [])]}
    with open(os.path.join(path, "config.json"), "w", encoding="utf-8") as f:
        json.dump(cfg, f)
    with open(os.path.join(path, "preprocessor_config.json"), "w", encoding="utf-8") as f:
        json.dump({"feature_size": dims.n_mels, "sampling_rate": 16000, "hop_length": 160, "chunk_length": 30, "n_fft": 400}, f)
    if tokenizer is not None:
        tokenizer.save(os.path.join(path, "tokenizer.json"))
