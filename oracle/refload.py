"""Import the UNMODIFIED reference package from /root/reference with its absent native deps stubbed.

TEST INFRASTRUCTURE ONLY (see oracle/whisper_oracle.py).  ``av``, ``ctranslate2`` and ``onnxruntime`` are
not installed in this image; the reference's pure-Python host layer imports fine once
empty modules of those names exist.  ``ctranslate2`` can instead be bound to a shim backed by an engine
(``oracle/ct2_shim.py``) so the reference's own ``transcribe.py`` drives our engine or the oracle.
Nothing here may be used on the GPU box: /root/reference does not exist there.
"""
import importlib
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "faster_whisper"))


def load_reference(ct2_module=None):
    """Returns the imported ``faster_whisper`` reference package (fresh import)."""
    if not reference_available():
        raise RuntimeError("reference tree not present")
    for name in list(sys.modules):
        if name == "faster_whisper" or name.startswith("faster_whisper."):
            del sys.modules[name]
    for name in ("av", "av.audio", "av.audio.fifo", "av.audio.resampler", "onnxruntime"):
        if name not in sys.modules:
            stub = types.ModuleType(name)
            # a real spec: importlib.util.find_spec (used by transformers' availability probes) raises on modules whose __spec__ is None
            stub.__spec__ = importlib.machinery.ModuleSpec(name, None)
            sys.modules[name] = stub
    if ct2_module is None:
        ct2_module = types.ModuleType("ctranslate2")
        ct2_module.models = types.ModuleType("ctranslate2.models")
        ct2_module.StorageView = type("StorageView", (), {})
        ct2_module.models.Whisper = type("Whisper", (), {})
        ct2_module.models.WhisperGenerationResult = type("WhisperGenerationResult", (), {})
    sys.modules["ctranslate2"] = ct2_module
    sys.modules["ctranslate2.models"] = ct2_module.models
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    try:
        return importlib.import_module("faster_whisper")
    finally:
        sys.path.remove(REFERENCE_ROOT)
