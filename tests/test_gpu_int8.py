"""-m gpu: compute_type="int8_float16" (BASELINE.json configs[3]) against the CPU oracle run on the SAME quantised weights.

The engine quantises every decoder linear layer and the output embedding per output channel (symmetric, scale = max|w| / 127) AFTER
folding the preceding LayerNorm's gamma into the weight, and both persistent step kernels stream the int8 values (dstep.cu for <= 8
rows, bstep.cu for more) with the scale applied to the fp32 accumulator.  `quantised_weights` restates that in NumPy and hands the
de-quantised, un-folded matrices to the fp32 oracle, so the comparison isolates the kernels from the quantisation noise:
tokens exact unless the oracle reports a near-tie, scores within 0.05.  (Reference: compute_type at faster_whisper/transcribe.py:689-698.)
"""
import numpy as np
import pytest

from faster_whisper_b200 import engine
from faster_whisper_b200.synthetic import synthetic_audio
from oracle import whisper_oracle as orc

pytestmark = pytest.mark.gpu
LOGIT_TOL = 0.05


def _quant_rows(wf: np.ndarray) -> np.ndarray:
    """csrc/dstep.cu ds_quant_rows_kernel: per row scale = max|w| / 127 (fp32), q = rint(w / scale) clamped to +-127; returns q * scale."""
    w16 = wf.astype(np.float16).astype(np.float32)
    mx = np.abs(w16).max(axis=1, keepdims=True)
    sc = np.where(mx > 0, mx / np.float32(127.0), np.float32(1.0)).astype(np.float32)
    q = np.clip(np.rint(w16 / sc), -127, 127).astype(np.float32)
    return q * sc


def quantised_weights(dims, w):
    out = dict(w)
    for i in range(dims.n_text_layer):
        p = f"decoder.blocks.{i}"
        for ln, names in ((".attn_ln", [".attn.query", ".attn.key", ".attn.value"]), (".cross_attn_ln", [".cross_attn.query"]), (".mlp_ln", [".mlp.0"])):
            g = w[p + ln + ".weight"].astype(np.float32)
            for n in names:
                out[p + n + ".weight"] = (_quant_rows(w[p + n + ".weight"] * g[None, :]) / g[None, :]).astype(np.float32)
        for n in (".attn.out", ".cross_attn.out", ".mlp.2"):
            out[p + n + ".weight"] = _quant_rows(w[p + n + ".weight"])
    gf = w["decoder.ln.weight"].astype(np.float32)
    out["decoder.output_projection.weight"] = (_quant_rows(w["decoder.token_embedding.weight"] * gf[None, :]) / gf[None, :]).astype(np.float32)
    return out


def features_for(m, n_chunks, seed=0):
    return np.stack([orc.pad_or_trim(orc.log_mel(synthetic_audio(seed + i, 30.0), m["dims"].n_mels)[:, :-1]) for i in range(n_chunks)])


@pytest.fixture(scope="module")
def int8_pair(micro_ml):
    dims, w, st = micro_ml["dims"], micro_ml["weights"], micro_ml["tokens"]
    eng = engine.Whisper(dims=dims, weights=w, tokens=st, device="cuda", compute_type="int8_float16")
    oracle_q = orc.WhisperOracle(dims.to_dict(), quantised_weights(dims, w), st.to_dict())
    return eng, oracle_q


@pytest.mark.parametrize("n_chunks,beam", [(1, 5), (3, 1), (4, 5), (16, 5)])  # 5 / 3 rows: dstep_kernel<W8>; 20 / 80 rows: bstep_kernel int8 atoms
def test_int8_decode_matches_oracle_on_quantised_weights(micro_ml, int8_pair, n_chunks, beam):
    eng, oq = int8_pair
    st = micro_ml["tokens"]
    feats = features_for(micro_ml, n_chunks, seed=400)
    prompts = [[st.sot, st.lang_begin, st.transcribe, st.no_timestamps]] * n_chunks
    kw = dict(beam_size=beam, max_length=36, repetition_penalty=1.25, no_repeat_ngram_size=3, return_scores=True, return_no_speech_prob=True)
    want = oq.generate(oq.encode(feats), prompts, **kw)
    got = eng.generate(eng.encode(feats), prompts, **kw)
    exact = 0
    for i, (w, g) in enumerate(zip(want, got)):
        assert abs(g.no_speech_prob - w.no_speech_prob) < 5e-3 * max(1.0, w.no_speech_prob) + 1e-6
        if g.sequences_ids[0] == w.sequences_ids[0]:
            exact += 1
            assert abs(g.scores[0] - w.scores[0]) < 0.05, (i, g.scores[0], w.scores[0])
        else:
            first = next((j for j, (x, y) in enumerate(zip(g.sequences_ids[0], w.sequences_ids[0])) if x != y), -1)
            print("chunk %d diverges at token %d, oracle min margin %.4f" % (i, first, w.min_margin))
            assert w.min_margin < 2 * LOGIT_TOL, (i, first, w.min_margin)
    assert exact >= 1  # every divergence above sits at a near-tie of the oracle (margins printed); greedy micro models have many


def test_int8_logits_close_to_quantised_oracle(micro_ml, int8_pair):
    """Teacher-forced logits of the int8 engine vs the oracle on the de-quantised weights: <= 0.05 (the prefill path multiplies the
    fp16-rounded de-quantised values, the oracle the exact q * scale)."""
    import torch

    eng, oq = int8_pair
    st = micro_ml["tokens"]
    feats = features_for(micro_ml, 2, seed=410)
    rng = np.random.default_rng(8)
    toks = np.concatenate([np.array([[st.sot, st.lang_begin, st.transcribe, st.no_timestamps]] * 2), rng.integers(0, 50000, (2, 12))], axis=1).astype(np.int32)
    enc_o = oq.encode(feats)
    want = oq.decoder_forward(torch.from_numpy(toks).long(), 0, [None] * micro_ml["dims"].n_text_layer, oq.cross_kv(enc_o), torch.arange(2)).numpy()
    got = eng.debug_logits(eng.encode(feats), toks)
    assert np.abs(got - want).max() < LOGIT_TOL, np.abs(got - want).max()
