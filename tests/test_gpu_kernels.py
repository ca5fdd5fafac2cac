"""-m gpu: each CUDA kernel of the hot path against the CPU oracle / NumPy, through the C ABI.

Tolerances are stated per test.  Inputs are seeded; sizes are kept where the oracle finishes in seconds.
"""
import numpy as np
import pytest

from faster_whisper_b200 import engine
from faster_whisper_b200.synthetic import synthetic_audio
from oracle import whisper_oracle as orc

pytestmark = pytest.mark.gpu


def f16(x):
    return np.asarray(x, np.float32).astype(np.float16).astype(np.float32)


# ---- log-mel -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("n", [0, 1, 159, 160, 161, 4000, 16000, 176000, 480000])
def test_logmel_matches_oracle(n_mels, n):
    x = synthetic_audio(5, n / 16000.0)
    want = orc.log_mel(x, n_mels)
    got = engine.log_mel(x, n_mels)
    assert got.shape == want.shape
    err = np.abs(got - want)
    # float32 DFT by direct summation vs pocketfft: agreement to ~1e-4 typical, 2e-3 at the clamp floor
    assert err.max() < 2e-3, (err.max(), err.mean())
    assert err.mean() < 1e-4, err.mean()


def test_logmel_silence_and_impulse():
    z = np.zeros(16000, np.float32)
    got = engine.log_mel(z, 80)
    assert np.allclose(got, -1.5)
    imp = z.copy()
    imp[8000] = 1.0
    got, want = engine.log_mel(imp, 80), orc.log_mel(imp, 80)
    assert np.abs(got - want).max() < 2e-3


# ---- dense GEMM (tcgen05) -----------------------------------------------------------------------------
@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("shape", [(128, 128, 64), (300, 384, 128), (1500, 1280, 1280), (3001, 256, 192), (257, 5120, 320)])
def test_gemm_vs_numpy(impl, shape):
    M, N, K = shape
    rng = np.random.default_rng(M * 7 + N)
    a, w, b = rng.standard_normal((M, K), np.float32), rng.standard_normal((N, K), np.float32) / np.sqrt(K), rng.standard_normal(N).astype(np.float32)
    want = f16(a) @ f16(w).T + b
    got = engine.debug_gemm(a, w, b, impl=impl)
    assert np.abs(got - want).max() < 2e-3 * max(1.0, np.abs(want).max()), np.abs(got - want).max()


def test_gemm_gelu_epilogue():
    import math

    rng = np.random.default_rng(3)
    a, w, b = rng.standard_normal((200, 256), np.float32), rng.standard_normal((384, 256), np.float32) / 16, rng.standard_normal(384).astype(np.float32)
    y = f16(a) @ f16(w).T + b
    want = 0.5 * y * (1 + np.vectorize(math.erf)(y / math.sqrt(2)))
    got = engine.debug_gemm(a, w, b, impl=0, gelu=True)
    assert np.abs(got - want).max() < 4e-3


# ---- encoder attention (tcgen05 flash) -------------------------------------------------------------------
@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("T", [1500, 128, 77])
def test_attention_vs_numpy(impl, T):
    rng = np.random.default_rng(T)
    B, H = 2, 2
    qkv = rng.standard_normal((B, T, 3 * H * 64), np.float32)
    got = engine.debug_attention(qkv, H, impl=impl)
    x = f16(qkv).reshape(B, T, 3, H, 64)
    q, k, v = x[:, :, 0], x[:, :, 1], x[:, :, 2]
    s = np.einsum("bqhd,bkhd->bhqk", q, k) / 8.0
    p = np.exp(s - s.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    want = np.einsum("bhqk,bkhd->bqhd", p, v).reshape(B, T, H * 64)
    # fp16 probabilities and fp16 output rounding
    assert np.abs(got - want).max() < 6e-3, np.abs(got - want).max()


# ---- decode skinny GEMM (mma.sync, weights streamed once) ------------------------------------------------
@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("R", [1, 5, 8, 17, 40, 80])
@pytest.mark.parametrize("NK", [(384, 384), (1280, 1280), (1000, 5120), (51864, 128)])
def test_skinny_gemm_vs_numpy(impl, R, NK):
    N, K = NK
    if impl == 1 and N * K * R > 3e8:
        pytest.skip("reference kernel is slow")
    rng = np.random.default_rng(R * 1000 + N)
    x, w, b = rng.standard_normal((R, K), np.float32), rng.standard_normal((N, K), np.float32) / np.sqrt(K), rng.standard_normal(N).astype(np.float32)
    want = f16(x) @ f16(w).T + b
    got = engine.debug_gemv(x, w, b, impl=impl)
    assert np.abs(got - want).max() < 2e-3 * max(1.0, np.abs(want).max())
