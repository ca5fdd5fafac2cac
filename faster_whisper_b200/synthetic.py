"""Seeded synthetic checkpoints and tokenizers at exact Whisper shapes.

No checkpoint or ``tokenizer.json`` exists offline, so the bench, the parity tests and
the CPU oracle all draw their weights from here: same names, same shapes, same float32 values on both
sides.  Names follow the OpenAI Whisper state dict (``encoder.blocks.0.attn.query.weight`` ...), which is
also what the CTranslate2 converter consumes.

Distribution ("sharpened" per SURVEY.md §7.4 so greedy margins are not dominated by fp16 noise):
linear weights N(0, 1/fan_in); biases N(0, 0.02^2); LayerNorm gamma 1+N(0,0.1^2), beta N(0,0.05^2);
sinusoidal encoder positions; decoder positions N(0, 0.01^2); tied token embedding N(0, (4/sqrt(d))^2)
so the final logits have a standard deviation of about 4.
"""

from __future__ import annotations

import zlib
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, Tuple

import numpy as np

from .config import LANGUAGE_CODES, MODEL_DIMS, WhisperDims, special_tokens


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> np.ndarray:
    """Whisper's fixed encoder position table (stored as a weight by the converter)."""
    half = channels // 2
    log_inc = np.log(max_timescale) / (half - 1)
    inv = np.exp(-log_inc * np.arange(half, dtype=np.float64))
    t = np.arange(length, dtype=np.float64)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)


def weight_specs(dims: WhisperDims) -> Dict[str, Tuple[Tuple[int, ...], str, float]]:
    """name -> (shape, kind, scale).  kind in {normal, ln_w, ln_b, sinusoid}."""
    d, da = dims.n_text_state, dims.n_audio_state
    specs: Dict[str, Tuple[Tuple[int, ...], str, float]] = {}

    def lin(prefix, out_f, in_f, bias=True):
        specs[prefix + ".weight"] = ((out_f, in_f), "normal", 1.0 / np.sqrt(in_f))
        if bias:
            specs[prefix + ".bias"] = ((out_f,), "normal", 0.02)

    def ln(prefix, n):
        specs[prefix + ".weight"] = ((n,), "ln_w", 0.1)
        specs[prefix + ".bias"] = ((n,), "ln_b", 0.05)

    def attn(prefix, n):
        lin(prefix + ".query", n, n)
        lin(prefix + ".key", n, n, bias=False)
        lin(prefix + ".value", n, n)
        lin(prefix + ".out", n, n)

    specs["encoder.conv1.weight"] = ((da, dims.n_mels, 3), "normal", 1.0 / np.sqrt(3 * dims.n_mels))
    specs["encoder.conv1.bias"] = ((da,), "normal", 0.02)
    specs["encoder.conv2.weight"] = ((da, da, 3), "normal", 1.0 / np.sqrt(3 * da))
    specs["encoder.conv2.bias"] = ((da,), "normal", 0.02)
    specs["encoder.positional_embedding"] = ((dims.n_audio_ctx, da), "sinusoid", 1.0)
    for i in range(dims.n_audio_layer):
        p = f"encoder.blocks.{i}"
        ln(p + ".attn_ln", da)
        attn(p + ".attn", da)
        ln(p + ".mlp_ln", da)
        lin(p + ".mlp.0", 4 * da, da)
        lin(p + ".mlp.2", da, 4 * da)
    ln("encoder.ln_post", da)

    specs["decoder.token_embedding.weight"] = ((dims.n_vocab, d), "normal", 4.0 / np.sqrt(d))
    specs["decoder.positional_embedding"] = ((dims.n_text_ctx, d), "normal", 0.01)
    for i in range(dims.n_text_layer):
        p = f"decoder.blocks.{i}"
        ln(p + ".attn_ln", d)
        attn(p + ".attn", d)
        ln(p + ".cross_attn_ln", d)
        attn(p + ".cross_attn", d)
        ln(p + ".mlp_ln", d)
        lin(p + ".mlp.0", 4 * d, d)
        lin(p + ".mlp.2", d, 4 * d)
    ln("decoder.ln", d)
    return specs


def _tensor(name: str, shape, kind: str, scale: float, seed: int) -> np.ndarray:
    rng = np.random.default_rng([seed, zlib.crc32(name.encode())])
    if kind == "sinusoid":
        return sinusoids(shape[0], shape[1])
    x = rng.standard_normal(size=shape, dtype=np.float32)
    x *= np.float32(scale)
    if kind == "ln_w":
        x += np.float32(1.0)
    return x


def make_weights(dims: WhisperDims | str, seed: int = 0, threads: int = 8) -> Dict[str, np.ndarray]:
    """All tensors of one synthetic checkpoint as float32 arrays (C-contiguous)."""
    if isinstance(dims, str):
        dims = MODEL_DIMS[dims]
    specs = weight_specs(dims)
    items = sorted(specs.items(), key=lambda kv: -int(np.prod(kv[1][0])))
    with ThreadPoolExecutor(max_workers=max(1, threads)) as pool:
        arrays = list(pool.map(lambda kv: _tensor(kv[0], *kv[1], seed), items))
    return {name: arr for (name, _), arr in zip(items, arrays)}


def custom_dims(
    name="micro", n_mels=80, d=128, heads=2, enc_layers=2, dec_layers=2, n_vocab=51864
) -> WhisperDims:
    """A shrunken geometry for fast CPU-side tests (head_dim stays 64)."""
    assert d % 64 == 0 and d // heads == 64, "head_dim must be 64 (all Whisper sizes)"
    return WhisperDims(name, n_mels, d, heads, enc_layers, d, heads, dec_layers, n_vocab)


def make_tokenizer(n_vocab: int):
    """A byte-level BPE ``tokenizers.Tokenizer`` whose control tokens sit at Whisper's ids.

    The text vocabulary is synthetic (256 byte symbols, two-letter merges, then unreachable filler
    entries), but every id the hot path cares about — eot, sot, languages, task tokens, no_speech,
    no_timestamps, timestamps — is where a real ``tokenizer.json`` puts it, and token 220 is " " as in
    GPT-2, which the reference's ``suppress_blank`` relies on.
    """
    from tokenizers import AddedToken, Tokenizer, decoders, models, pre_tokenizers

    st = special_tokens(n_vocab)
    alphabet = sorted(pre_tokenizers.ByteLevel.alphabet())
    vocab = {ch: i for i, ch in enumerate(alphabet)}
    assert alphabet[220] == "Ġ"
    letters = "abcdefghijklmnopqrstuvwxyz"
    merges = []
    for a in ["Ġ"] + list(letters):
        for b in letters:
            if len(vocab) >= st.eot:
                break
            vocab[a + b] = len(vocab)
            merges.append((a, b))
    k = 0
    while len(vocab) < st.eot:
        vocab[f"Ġw{k:05d}q"] = len(vocab)
        k += 1
    tok = Tokenizer(models.BPE(vocab=vocab, merges=merges))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tok.decoder = decoders.ByteLevel()
    names = ["<|endoftext|>", "<|startoftranscript|>"]
    names += [f"<|{c}|>" for c in (LANGUAGE_CODES + ["xx1", "xx2", "xx3"])[: st.num_languages]]
    names += ["<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>"]
    names += ["<|nospeech|>" if n_vocab == 51866 else "<|nocaptions|>", "<|notimestamps|>"]
    names += [f"<|{i * 0.02:.2f}|>" for i in range(1501)]
    tok.add_special_tokens([AddedToken(n, special=True) for n in names])
    assert tok.token_to_id("<|endoftext|>") == st.eot
    assert tok.token_to_id("<|notimestamps|>") == st.no_timestamps
    assert tok.get_vocab_size() == n_vocab, (tok.get_vocab_size(), n_vocab)
    return tok


def synthetic_audio(index: int, seconds: float = 30.0, sampling_rate: int = 16000) -> np.ndarray:
    """Chunk *index* of the benchmark workload: 0.1*N(0,1) noise plus a slow chirp so
    the log-mel is not flat.  float32, deterministic."""
    n = int(round(seconds * sampling_rate))
    rng = np.random.default_rng(1000 + index)
    x = 0.1 * rng.standard_normal(n, dtype=np.float32)
    t = np.arange(n, dtype=np.float32) / np.float32(sampling_rate)
    f0, f1 = 200.0 + 37.0 * (index % 7), 4000.0
    phase = 2 * np.pi * (f0 * t + (f1 - f0) * t * t / (2 * np.float32(max(seconds, 1e-3))))
    x += (0.05 * np.sin(phase)).astype(np.float32)
    return x.astype(np.float32)
