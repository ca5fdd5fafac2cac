"""CPU: the Hugging Face checkpoint reader (faster_whisper_b200/hf_format.py) against a directory written by
``transformers.WhisperForConditionalGeneration.save_pretrained`` — an independent writer of both the safetensors container
and the parameter naming (reference faster_whisper/utils.py:91-97, transcribe.py:700-710 load model
directories).  Every tensor must come back bit for bit under the OpenAI-Whisper name the engine consumes."""
import io
import os

import numpy as np
import pytest

from faster_whisper_b200.checkpoint import load_model_dir
from faster_whisper_b200.hf_format import hf_to_openai_name, is_hf_dir, load_hf_dir, read_safetensors
from faster_whisper_b200.synthetic import custom_dims, make_weights


@pytest.fixture(scope="module")
def hf_dir(tmp_path_factory):
    pytest.importorskip("transformers")
    import torch
    from transformers import WhisperConfig, WhisperForConditionalGeneration

    from oracle.check_against_transformers import to_hf_state

    dims = custom_dims(d=128, heads=2, enc_layers=2, dec_layers=3, n_vocab=51864)
    w = make_weights(dims, seed=21)
    cfg = WhisperConfig(vocab_size=dims.n_vocab, num_mel_bins=dims.n_mels, d_model=dims.n_text_state, encoder_layers=dims.n_audio_layer,
                        decoder_layers=dims.n_text_layer, encoder_attention_heads=dims.n_audio_head, decoder_attention_heads=dims.n_text_head,
                        encoder_ffn_dim=4 * dims.n_audio_state, decoder_ffn_dim=4 * dims.n_text_state, max_source_positions=1500,
                        max_target_positions=448, activation_function="gelu")
    hf = WhisperForConditionalGeneration(cfg)
    missing, unexpected = hf.load_state_dict(to_hf_state(w, dims), strict=False)
    assert not unexpected
    path = str(tmp_path_factory.mktemp("hf_whisper"))
    hf.save_pretrained(path, safe_serialization=True)
    zero_k_bias = {k for k in (m.replace("model.", "") for m in missing)}
    return path, dims, w, hf, zero_k_bias


def test_hf_directory_round_trips_every_tensor(hf_dir):
    path, dims, w, hf, _ = hf_dir
    assert is_hf_dir(path)
    got_dims, got, cfg = load_hf_dir(path)
    assert (got_dims.n_mels, got_dims.n_audio_state, got_dims.n_audio_head, got_dims.n_audio_layer, got_dims.n_text_layer, got_dims.n_vocab) == (
        dims.n_mels, dims.n_audio_state, dims.n_audio_head, dims.n_audio_layer, dims.n_text_layer, dims.n_vocab)
    for name, arr in w.items():
        assert name in got, name
        assert got[name].shape == arr.shape and np.array_equal(got[name].astype(np.float32), arr), name
    extra = set(got) - set(w)
    # transformers materialises the (absent) key-projection bias as zeros; nothing else may appear
    assert all(n.endswith("key.bias") and not got[n].any() for n in extra), extra
    # the generic entry point finds the format by itself, also from in-memory files (file-like entries are not consumed by probing)
    d2, w2 = load_model_dir(path)
    assert d2.n_text_layer == dims.n_text_layer and np.array_equal(w2["decoder.blocks.2.mlp.2.weight"], w["decoder.blocks.2.mlp.2.weight"])
    files = {n: io.BytesIO(open(os.path.join(path, n), "rb").read()) for n in ("config.json", "model.safetensors")}
    d3, w3 = load_model_dir("", files)
    assert np.array_equal(w3["encoder.conv1.weight"], w["encoder.conv1.weight"])


def test_safetensors_parser_dtypes_and_errors(tmp_path):
    pytest.importorskip("safetensors")
    import torch
    from safetensors.torch import save_file

    t = {"a": torch.arange(12, dtype=torch.float32).reshape(3, 4), "b": torch.arange(6, dtype=torch.float16), "c": torch.tensor([1.5, -2.25], dtype=torch.bfloat16)}
    p = str(tmp_path / "x.safetensors")
    save_file(t, p)
    got = read_safetensors(p)
    assert np.array_equal(got["a"], t["a"].numpy()) and got["b"].dtype == np.float16
    assert got["c"].dtype == np.float32 and got["c"].tolist() == [1.5, -2.25]
    with pytest.raises(ValueError):
        read_safetensors(b"\x00\x01")
    assert hf_to_openai_name("model.decoder.layers.3.encoder_attn.q_proj.weight") == "decoder.blocks.3.cross_attn.query.weight"
    assert hf_to_openai_name("proj_out.weight") is None
