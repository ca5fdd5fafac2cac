"""A/B of the many-row step kernel's wave count in ONE process: large-v3 (synthetic), B chunks x beam 5, prompt 4 + N new tokens.

For each value of B2W_BSTEP_WAVES (re-read by the engine on every generate call) one warm-up and `--repeat` timed generate calls;
prints the decode time per step from the engine's stage timers (CUDA events on its stream) and checks that every wave count
produces the same tokens as the first one.  With --prof the per-phase device timers of the last step are printed too.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="large-v3")
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--new-tokens", type=int, default=128)
ap.add_argument("--beam", type=int, default=5)
ap.add_argument("--repeat", type=int, default=2)
ap.add_argument("--waves", default="1,2,3,4", help="comma list; an entry may carry a prefetch-gate mask as <waves>g<mask>, e.g. 1g3")
ap.add_argument("--compute-type", default="float16")
ap.add_argument("--prof", action="store_true")
a = ap.parse_args()
if a.prof:
    os.environ["B2W_DSTEP_PROF"] = "1"

from faster_whisper_b200 import engine  # noqa: E402
from faster_whisper_b200.config import MODEL_DIMS, special_tokens  # noqa: E402
from faster_whisper_b200.synthetic import make_weights, synthetic_audio  # noqa: E402

dims = MODEL_DIMS[a.model]
st = special_tokens(dims.n_vocab)
eng = engine.Whisper(dims=dims, weights=make_weights(dims, seed=0), tokens=st, device="cuda", compute_type=a.compute_type)
chunks = [synthetic_audio(i, 30.0) for i in range(a.batch)]
prompt = [st.sot, st.lang_begin, st.transcribe, st.no_timestamps] if dims.is_multilingual else [st.sot, st.no_timestamps]
sup = sorted({st.eot, st.sot, st.transcribe, st.translate, st.sot_prev, st.sot_lm, st.no_speech})
eng.timing(enable=True)
enc = eng.encode_audio(chunks)
eng.sync()
first = None
for waves in [w for w in a.waves.split(",") if w]:
    os.environ["B2W_BSTEP_WAVES"] = waves.split("g")[0]
    if "g" in waves:
        os.environ["B2W_BSTEP_GATE"] = waves.split("g")[1]
    else:
        os.environ.pop("B2W_BSTEP_GATE", None)
    best = None
    for it in range(1 + a.repeat):
        eng.timing(reset=True)
        res = eng.generate(enc, [prompt] * a.batch, beam_size=a.beam, max_length=len(prompt) + a.new_tokens, suppress_tokens=sup, return_scores=True)
        eng.sync()
        t = eng.timing()
        if it > 0:
            ms = t["decode_ms"] / max(1, t["decode_steps"])
            best = ms if best is None else min(best, ms)
    toks = [r.sequences_ids[0] for r in res]
    if first is None:
        first = toks
    same = sum(x == y for x, y in zip(first, toks))
    print("%swaves %s: decode %.4f ms/step (%d steps), tokens identical to the first setting in %d of %d chunks" % (os.environ.get("B2W_LIBRARY", "") and "[other build] ", waves, best, t["decode_steps"], same, len(toks)), flush=True)
