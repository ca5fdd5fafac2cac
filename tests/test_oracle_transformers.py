"""CPU: the oracle's network math (encoder, teacher-forced decoder, incremental KV path) against
``transformers.WhisperForConditionalGeneration`` carrying the same synthetic weights.

transformers is NOT the reference (CTranslate2 is, and it cannot be installed here — SURVEY.md §8c): this pins our
restatement of the Whisper architecture against an independent implementation, nothing more.  Tolerances are ~10x the
differences measured with torch 2.11 CPU fp32 (encoder 2.3e-6, logits 1.0e-5, incremental-vs-parallel 1.2e-5).
"""
import pytest


def test_oracle_network_math_matches_transformers():
    pytest.importorskip("transformers")
    from oracle.check_against_transformers import run_check

    enc_err, logit_err, inc_err = run_check()
    assert enc_err < 2e-5, enc_err
    assert logit_err < 1e-4, logit_err
    assert inc_err < 1e-4, inc_err


def _micro_oracle():
    from faster_whisper_b200.config import special_tokens
    from faster_whisper_b200.synthetic import custom_dims, make_weights
    from oracle.whisper_oracle import WhisperOracle

    dims = custom_dims(d=128, heads=2, enc_layers=1, dec_layers=1, n_vocab=51866)
    st = special_tokens(dims.n_vocab)
    return dims, st, WhisperOracle(dims.to_dict(), make_weights(dims, seed=1), st.to_dict())


def test_timestamp_rules_match_transformers_processor():
    """The oracle's Whisper timestamp rules (restated from CTranslate2's contract = OpenAI's ApplyTimestampRules) against
    transformers' independent ``WhisperTimeStampLogitsProcessor``: identical sets of allowed tokens at every step of 300 random
    decodes (histories are extended with the oracle's own greedy choice, so only reachable states are visited; timestamp mass is
    boosted in half of the steps so that the "timestamp mass beats every text token" rule fires)."""
    pytest.importorskip("transformers")
    import types

    import numpy as np
    import torch
    from transformers.generation.logits_process import WhisperTimeStampLogitsProcessor

    from oracle.whisper_oracle import LOWEST

    dims, st, orc = _micro_oracle()
    V = dims.n_vocab
    cfg = types.SimpleNamespace(eos_token_id=st.eot, no_timestamps_token_id=st.no_timestamps, max_initial_timestamp_index=50, forced_decoder_ids=None)
    prompt = [st.sot, st.lang_begin, st.transcribe]
    proc = WhisperTimeStampLogitsProcessor(cfg, begin_index=len(prompt))
    o = dict(repetition_penalty=1.0, no_repeat_ngram_size=0, suppress_ids=np.zeros(0, np.int64), suppress_blank=False, timestamp_rules=True,
             max_initial_timestamp_index=50)
    rng = np.random.default_rng(0)
    steps = forced = 0
    for _ in range(300):
        hist = []
        for step in range(int(rng.integers(1, 14))):
            logits = (rng.standard_normal((1, V)) * 3).astype(np.float32)
            if rng.random() < 0.5:
                logits[0, st.timestamp_begin:] += rng.uniform(0, 6)
            ours = logits.copy()
            orc._process_logits(ours, step, [hist], o)
            theirs = proc(torch.tensor([prompt + hist]), torch.from_numpy(logits.copy()))[0]
            allowed = ours[0] > LOWEST / 2
            assert np.array_equal(allowed, torch.isfinite(theirs).numpy()), (step, hist)
            steps += 1
            forced += int(step > 0 and not allowed[: st.timestamp_begin].any())
            t = int(np.argmax(ours[0]))
            if t == st.eot:
                break
            hist.append(t)
    assert steps > 1500 and forced > 50  # the forcing rules were exercised


def test_penalty_ngram_and_suppression_match_transformers_processors():
    """Repetition penalty, no-repeat-n-gram blocking and token suppression against transformers' processors of the same names
    (CTranslate2 documents both penalties by reference to them): same masked set, same penalised values."""
    pytest.importorskip("transformers")
    import numpy as np
    import torch
    from transformers.generation.logits_process import NoRepeatNGramLogitsProcessor, RepetitionPenaltyLogitsProcessor, SuppressTokensLogitsProcessor

    from oracle.whisper_oracle import LOWEST

    dims, st, orc = _micro_oracle()
    V = dims.n_vocab
    rng = np.random.default_rng(1)
    for _ in range(200):
        rp, ng = float(rng.choice([1.0, 1.1, 1.3, 2.0])), int(rng.choice([0, 1, 2, 3]))
        sup = sorted({int(x) for x in rng.integers(0, V, size=5)})
        hist = [int(x) for x in rng.integers(0, 30, size=int(rng.integers(0, 20)))]  # a small alphabet: repeats and n-gram hits
        logits = (rng.standard_normal((1, V)) * 3).astype(np.float32)
        ours = logits.copy()
        orc._process_logits(ours, len(hist), [hist], dict(repetition_penalty=rp, no_repeat_ngram_size=ng, suppress_ids=np.asarray(sup, np.int64),
                                                         suppress_blank=False, timestamp_rules=False, max_initial_timestamp_index=None))
        theirs, ids = torch.from_numpy(logits.copy()), torch.tensor([hist], dtype=torch.long)
        if rp != 1.0 and hist:
            theirs = RepetitionPenaltyLogitsProcessor(rp)(ids, theirs)
        if ng > 0 and hist:
            theirs = NoRepeatNGramLogitsProcessor(ng)(ids, theirs)
        theirs = SuppressTokensLogitsProcessor(sup)(ids, theirs).numpy()[0]
        allowed = ours[0] > LOWEST / 2
        assert np.array_equal(allowed, np.isfinite(theirs)), (rp, ng, hist)
        assert np.allclose(ours[0][allowed], theirs[allowed], rtol=0, atol=1e-6)


def test_greedy_decode_matches_transformers_generate_token_for_token():
    """The oracle's greedy loop (incremental KV cache, argmax, EOS handling, score bookkeeping) against transformers' own
    ``GenerationMixin.generate`` on the same synthetic weights and features: 40 new tokens for three chunks, identical — unless the
    oracle's own top-1 / top-2 gap at the step of divergence is below 1e-3 (fp32 summation order), which is printed."""
    pytest.importorskip("transformers")
    import warnings

    import numpy as np
    import torch
    from transformers import GenerationConfig, WhisperConfig, WhisperForConditionalGeneration
    from transformers.generation.utils import GenerationMixin

    from faster_whisper_b200.config import special_tokens
    from faster_whisper_b200.synthetic import custom_dims, make_weights
    from oracle.check_against_transformers import to_hf_state
    from oracle.whisper_oracle import WhisperOracle

    dims = custom_dims(d=128, heads=2, enc_layers=2, dec_layers=2, n_vocab=51864)
    w = make_weights(dims, seed=3)
    cfg = WhisperConfig(vocab_size=dims.n_vocab, num_mel_bins=dims.n_mels, d_model=dims.n_text_state, encoder_layers=2, decoder_layers=2,
                        encoder_attention_heads=2, decoder_attention_heads=2, encoder_ffn_dim=512, decoder_ffn_dim=512,
                        max_source_positions=1500, max_target_positions=448, activation_function="gelu", dropout=0.0,
                        attention_dropout=0.0, activation_dropout=0.0)
    hf = WhisperForConditionalGeneration(cfg).eval()
    hf.load_state_dict(to_hf_state(w, dims), strict=False)
    st = special_tokens(dims.n_vocab)
    orc = WhisperOracle(dims.to_dict(), w, st.to_dict())
    feats = np.random.default_rng(0).standard_normal((3, dims.n_mels, 3000), dtype=np.float32) * 0.5
    prompt, n_new = [st.sot, st.no_timestamps], 40
    gc = GenerationConfig(max_new_tokens=n_new, do_sample=False, num_beams=1, eos_token_id=st.eot, pad_token_id=st.eot, bos_token_id=st.sot,
                          decoder_start_token_id=st.sot)
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = GenerationMixin.generate(hf, encoder_outputs=hf.model.encoder(torch.from_numpy(feats)), decoder_input_ids=torch.tensor([prompt] * 3),
                                       generation_config=gc)
    res = orc.generate(orc.encode(feats), [prompt] * 3, beam_size=1, max_length=len(prompt) + n_new, suppress_blank=False, suppress_tokens=[])
    for b in range(3):
        theirs = out[b, len(prompt):].tolist()
        theirs = theirs[: theirs.index(st.eot)] if st.eot in theirs else theirs
        ours = res[b].sequences_ids[0]
        first = next((i for i, (x, y) in enumerate(zip(theirs, ours)) if x != y), None)
        if first is not None:
            print("chunk %d parts from transformers at token %d, oracle margin there %.2e" % (b, first, res[b].step_margins[first]))
            assert res[b].step_margins[first] < 1e-3
        else:
            assert len(theirs) == len(ours) >= 30
