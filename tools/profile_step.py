"""One short pass of the hot path for ncu: large-v3 (synthetic), B chunks, beam 5, a few new tokens."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from faster_whisper_b200 import engine
from faster_whisper_b200.config import MODEL_DIMS, special_tokens
from faster_whisper_b200.synthetic import make_weights, synthetic_audio

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="large-v3")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--new-tokens", type=int, default=6)
ap.add_argument("--beam", type=int, default=5)
ap.add_argument("--repeat", type=int, default=1)
a = ap.parse_args()
dims = MODEL_DIMS[a.model]
st = special_tokens(dims.n_vocab)
eng = engine.Whisper(dims=dims, weights=make_weights(dims, seed=0), tokens=st, device="cuda")
chunks = [synthetic_audio(i, 30.0) for i in range(a.batch)]
prompt = [st.sot, st.lang_begin, st.transcribe, st.no_timestamps] if dims.is_multilingual else [st.sot, st.no_timestamps]
sup = sorted({st.eot, st.sot, st.transcribe, st.translate, st.sot_prev, st.sot_lm, st.no_speech})
for _ in range(a.repeat):
    enc = eng.encode_audio(chunks)
    res = eng.generate(enc, [prompt] * a.batch, beam_size=a.beam, max_length=len(prompt) + a.new_tokens, suppress_tokens=sup, return_scores=True)
eng.sync()
print("tokens", res[0].sequences_ids[0], eng.timing())
