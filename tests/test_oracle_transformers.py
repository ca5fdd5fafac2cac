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
