"""CPU: the codec-token -> waveform oracle (oracle/code2wav_ref.py) against the golden vectors generated from
transformers' Qwen3OmniMoeCode2Wav (tests/golden/make_golden.py code2wav) -- the published cousin of the Qwen3-TTS 12 Hz
codec decoder (parity with the real, absent upstream stays unpinned)."""
import os

import numpy as np
import pytest

from oracle import code2wav_ref as C


@pytest.fixture(scope="module")
def case(golden_dir):
    G = np.load(os.path.join(golden_dir, "code2wav_micro.npz"))
    g = C.GEOMETRIES["micro"]
    return g, C.make_weights(g, 0), G


def test_pre_transformer_matches_transformers(case):
    g, w, G = case
    h = C.pre_transformer(w, g, G["codes"])
    assert h.shape == G["hidden"].shape and np.abs(h - G["hidden"]).max() < 2e-5


def test_waveform_matches_transformers(case):
    g, w, G = case
    wav = C.code2wav_forward(w, g, G["codes"])
    assert wav.shape == G["wav"].shape and wav.dtype == np.float32
    assert np.abs(G["wav"]).max() < 1.0          # the case is not saturated by the final clamp
    assert np.abs(wav - G["wav"]).max() < 2e-5


def test_chunked_streaming_decode_matches_transformers(case):
    g, w, G = case
    out = C.chunked_decode(w, g, G["codes"], int(G["chunk_size"]), int(G["left_context"]))
    assert out.shape == G["chunked"].shape and np.abs(out - G["chunked"]).max() < 2e-5


def test_length_and_frame_rate(case):
    """T frames -> T * total_upsample samples minus the trims of the decoder's transposed convolutions (kernel 2 x stride:
    one stride is trimmed per side; the kernel = stride ones of the upsample stage trim nothing): 24 samples per frame in
    the micro geometry, 1920 (= 12.5 Hz at 24 kHz) in the published one."""
    g, w, G = case
    T = G["codes"].shape[1]
    assert g.total_upsample == 24 and C.GEOMETRIES["qwen3-12hz"].total_upsample == 1920
    n = T
    for f in g.upsampling_ratios:
        n = n * f                                  # k = stride: (n - 1) * s + s
    for f in g.upsample_rates:
        n = n * f - f                              # k = 2 * stride: (n - 1) * s + 2s, minus s per side
    assert len(G["wav"]) == n


def test_future_frames_do_not_change_the_past(case):
    """The stack is causal up to one frame of look-ahead per transposed convolution: changing the last code frame leaves the
    samples of all but the last few frames untouched."""
    g, w, G = case
    codes = G["codes"].copy()
    base = C.code2wav_forward(w, g, codes)
    codes[:, -1] = (codes[:, -1] + 7) % g.codebook_size
    other = C.code2wav_forward(w, g, codes)
    keep = (codes.shape[1] - 1 - len(g.upsampling_ratios + g.upsample_rates)) * g.total_upsample - g.total_upsample
    assert np.array_equal(base[:keep], other[:keep]) and not np.array_equal(base, other)


def test_building_blocks_shapes():
    x = np.random.default_rng(0).standard_normal((6, 11)).astype(np.float32)
    w = np.random.default_rng(1).standard_normal((4, 6, 7)).astype(np.float32)
    assert C.causal_conv1d(x, w, np.zeros(4, np.float32), dilation=3).shape == (4, 11)          # causal: same length
    wt = np.random.default_rng(2).standard_normal((6, 3, 10)).astype(np.float32)
    assert C.causal_trans_conv1d(x, wt, np.zeros(3, np.float32), 5).shape == (3, 11 * 5 - 5)
    a = np.zeros(6, np.float32)
    assert np.allclose(C.snake_beta(x, a, a), x + np.sin(x) ** 2, atol=1e-6)
