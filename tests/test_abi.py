"""CPU: the C-ABI library builds, loads, exports every symbol the header declares, and fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

from faster_whisper_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200whisper.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2w_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    declared = header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/b200whisper.h but not exported"
    assert set(declared) == set(engine.ABI_SYMBOLS)
    assert lib.b2w_abi_version() == 1


def test_struct_layout_matches_header():
    # b2w_gen_opts: the ctypes mirror must have the size the C side fills in
    lib = engine.load_library()
    o = engine._GenOpts()
    ctypes.memset(ctypes.byref(o), 0xFF, ctypes.sizeof(o))
    lib.b2w_gen_opts_default(ctypes.byref(o))
    assert (o.beam_size, o.num_hypotheses, o.max_length, o.max_initial_timestamp_index, o.suppress_blank, o.sampling_topk) == (5, 1, 448, 50, 1, 1)
    assert (o.patience, o.length_penalty, o.repetition_penalty, o.sampling_temperature) == (1.0, 1.0, 1.0, 1.0)
    assert o.debug_fake_logits == 0 and o.seed == 0 and o.n_suppress_tokens == 0


def test_no_cpu_fallback():
    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device|CPU path"):
        engine.log_mel(np.zeros(1600, np.float32), 80)
    from faster_whisper_b200.synthetic import custom_dims

    dims = custom_dims()
    with pytest.raises((RuntimeError, ValueError)):
        engine.Whisper(dims=dims, weights={}, device="cuda")
    with pytest.raises(ValueError):
        engine.Whisper(dims=dims, weights={}, device="cpu")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "faster_whisper_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f
                assert not re.search(r"#include\s+[\"<].*oracle", src), f
