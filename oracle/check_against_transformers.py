"""Cross-check the oracle's *network math* against ``transformers.WhisperForConditionalGeneration``.

TEST INFRASTRUCTURE ONLY.  transformers is not the reference (CTranslate2 is, and it is absent — SURVEY.md
§8c); this only guards our restatement of the Whisper architecture against transcription mistakes.  Run:

    python oracle/check_against_transformers.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from faster_whisper_b200.config import special_tokens  # noqa: E402
from faster_whisper_b200.synthetic import custom_dims, make_weights  # noqa: E402
from oracle.whisper_oracle import WhisperOracle  # noqa: E402


def to_hf_state(weights, dims):
    sd = {}

    def put(hf, name):
        sd[hf] = torch.from_numpy(weights[name])

    put("model.encoder.conv1.weight", "encoder.conv1.weight")
    put("model.encoder.conv1.bias", "encoder.conv1.bias")
    put("model.encoder.conv2.weight", "encoder.conv2.weight")
    put("model.encoder.conv2.bias", "encoder.conv2.bias")
    put("model.encoder.embed_positions.weight", "encoder.positional_embedding")
    amap = {"query": "q_proj", "key": "k_proj", "value": "v_proj", "out": "out_proj"}

    def block(hf, name, cross):
        for ours, theirs in amap.items():
            put(f"{hf}.self_attn.{theirs}.weight", f"{name}.attn.{ours}.weight")
            if ours != "key":
                put(f"{hf}.self_attn.{theirs}.bias", f"{name}.attn.{ours}.bias")
        put(f"{hf}.self_attn_layer_norm.weight", f"{name}.attn_ln.weight")
        put(f"{hf}.self_attn_layer_norm.bias", f"{name}.attn_ln.bias")
        if cross:
            for ours, theirs in amap.items():
                put(f"{hf}.encoder_attn.{theirs}.weight", f"{name}.cross_attn.{ours}.weight")
                if ours != "key":
                    put(f"{hf}.encoder_attn.{theirs}.bias", f"{name}.cross_attn.{ours}.bias")
            put(f"{hf}.encoder_attn_layer_norm.weight", f"{name}.cross_attn_ln.weight")
            put(f"{hf}.encoder_attn_layer_norm.bias", f"{name}.cross_attn_ln.bias")
        put(f"{hf}.fc1.weight", f"{name}.mlp.0.weight")
        put(f"{hf}.fc1.bias", f"{name}.mlp.0.bias")
        put(f"{hf}.fc2.weight", f"{name}.mlp.2.weight")
        put(f"{hf}.fc2.bias", f"{name}.mlp.2.bias")
        put(f"{hf}.final_layer_norm.weight", f"{name}.mlp_ln.weight")
        put(f"{hf}.final_layer_norm.bias", f"{name}.mlp_ln.bias")

    for i in range(dims.n_audio_layer):
        block(f"model.encoder.layers.{i}", f"encoder.blocks.{i}", False)
    put("model.encoder.layer_norm.weight", "encoder.ln_post.weight")
    put("model.encoder.layer_norm.bias", "encoder.ln_post.bias")
    put("model.decoder.embed_tokens.weight", "decoder.token_embedding.weight")
    put("model.decoder.embed_positions.weight", "decoder.positional_embedding")
    for i in range(dims.n_text_layer):
        block(f"model.decoder.layers.{i}", f"decoder.blocks.{i}", True)
    put("model.decoder.layer_norm.weight", "decoder.ln.weight")
    put("model.decoder.layer_norm.bias", "decoder.ln.bias")
    put("proj_out.weight", "decoder.token_embedding.weight")
    return sd


def run_check():
    """Returns (encoder, logits, incremental-vs-parallel) max abs differences; tests/test_oracle_transformers.py asserts on them."""
    from transformers import WhisperConfig, WhisperForConditionalGeneration

    dims = custom_dims(d=128, heads=2, enc_layers=2, dec_layers=2, n_vocab=51864)
    w = make_weights(dims, seed=3)
    cfg = WhisperConfig(
        vocab_size=dims.n_vocab, num_mel_bins=dims.n_mels, d_model=dims.n_text_state,
        encoder_layers=dims.n_audio_layer, decoder_layers=dims.n_text_layer,
        encoder_attention_heads=dims.n_audio_head, decoder_attention_heads=dims.n_text_head,
        encoder_ffn_dim=4 * dims.n_audio_state, decoder_ffn_dim=4 * dims.n_text_state,
        max_source_positions=1500, max_target_positions=448, activation_function="gelu",
        dropout=0.0, attention_dropout=0.0, activation_dropout=0.0,
    )
    hf = WhisperForConditionalGeneration(cfg).eval()
    missing, unexpected = hf.load_state_dict(to_hf_state(w, dims), strict=False)
    assert not unexpected, unexpected
    assert all("k_proj.bias" in m for m in missing) or not missing, missing
    st = special_tokens(dims.n_vocab)
    orc = WhisperOracle(dims.to_dict(), w, st.to_dict())
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((2, dims.n_mels, 3000), dtype=np.float32) * 0.5
    toks = torch.tensor([[st.sot, st.no_timestamps, 11, 22, 33], [st.sot, st.no_timestamps, 44, 55, 66]])
    with torch.no_grad():
        enc_hf = hf.model.encoder(torch.from_numpy(feats)).last_hidden_state
        out_hf = hf(input_features=torch.from_numpy(feats), decoder_input_ids=toks).logits
    enc = orc.encode(feats)
    xkv = orc.cross_kv(enc)
    cache = [None] * dims.n_text_layer
    logits = orc.decoder_forward(toks, 0, cache, xkv, torch.arange(2))
    # incremental path must agree with the parallel one
    cache2 = [None] * dims.n_text_layer
    inc = [orc.decoder_forward(toks[:, i : i + 1], i, cache2, xkv, torch.arange(2)) for i in range(toks.shape[1])]
    inc = torch.cat(inc, dim=1)
    e1 = float((enc - enc_hf).abs().max())
    e2 = float((logits - out_hf).abs().max())
    e3 = float((inc - logits).abs().max())
    return e1, e2, e3


def main():
    e1, e2, e3 = run_check()
    print(f"encoder max|diff| {e1:.3e}  logits max|diff| {e2:.3e}  incremental-vs-parallel {e3:.3e}")
    # measured here (torch 2.11 CPU fp32): 2.3e-6 / 1.0e-5 / 1.2e-5
    assert e1 < 2e-5 and e2 < 1e-4 and e3 < 1e-4
    print("OK")


if __name__ == "__main__":
    main()
