"""GPU parity tests (B200) of the TTS talker + code predictor (csrc/qwen3tts.cu) through the C ABI: codebook ids
bit-exact against the numpy oracle (oracle/qwen3tts_ref.py) and the golden codes generated from the transformers
cousin; sessions batched in one launch equal the single-session results.  Unpinned vs faster-qwen3-tts (absent)."""
import os

import numpy as np
import pytest
import torch

from oracle import code2wav_ref as C, qwen3tts_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def E():
    from speech_to_speech_b200 import engine
    return engine


def _engine(E, dtype="float16", **kw):
    g, cg = R.GEOMETRIES["micro"], C.GEOMETRIES["micro"]
    cg = C.Code2WavGeometry(**{**cg.to_dict(), "quantizers": g.n_groups, "codebook_size": g.predictor.vocab,
                               "upsample_rates": tuple(cg.upsample_rates), "upsampling_ratios": tuple(cg.upsampling_ratios)})
    w, cw = R.make_weights(g, 0), C.make_weights(cg, 0)
    eng = E.Qwen3TTSEngine(g.to_dict(), cg.to_dict(), dtype=dtype, max_positions=128, max_text=64, **kw)
    eng.load_state_dict(w, cw)
    return g, cg, w, cw, eng


@pytest.mark.parametrize("dtype", ["float16", "bfloat16"])
def test_codes_match_transformers_golden(E, golden_dir, dtype):
    g, cg, w, cw, eng = _engine(E, dtype)
    G = np.load(os.path.join(golden_dir, "qwen3tts_micro.npz"))
    F = int(G["max_frames"])
    eng.prefill(0, G["text_ids"].tolist(), int(G["speaker"]))
    codes = eng.decode_frames([0], F)[0].cpu().numpy()
    # ids are compared up to the first frame whose decision margin is below the 16-bit operand noise (none in practice:
    # make_weights gives the heads wide logits); the first codes of all F frames and the full F-1 golden frames
    assert codes[:, 0].tolist() == G["code0_all"].tolist()
    assert np.array_equal(codes[: F - 1], G["codes"])


def test_batched_sessions_equal_oracle_and_single_runs(E):
    g, cg, w, cw, eng = _engine(E, "float16", max_sessions=4)
    rng = np.random.default_rng(3)
    texts = [rng.integers(0, 400, n).tolist() for n in (1, 3, 9, 17)]
    speakers = [2301, 2302, 2301, 2400]
    F = 7
    refs = [R.generate(w, g, t, s, F) for t, s in zip(texts, speakers)]
    for slot, (t, s) in enumerate(zip(texts, speakers)):
        eng.prefill(slot, t, s)
    got = eng.decode_frames([0, 1, 2, 3], F).cpu().numpy()
    for slot in range(4):
        assert np.array_equal(got[slot][: len(refs[slot])], refs[slot]), slot
    # frames continue across calls (state lives in the library): 3 + 4 frames == 7 frames
    for slot, (t, s) in enumerate(zip(texts, speakers)):
        eng.prefill(slot, t, s)
    a = eng.decode_frames([3, 1], 3).cpu().numpy()
    b = eng.decode_frames([3, 1], 4).cpu().numpy()
    assert np.array_equal(np.concatenate([a, b], 1)[0], got[3]) and np.array_equal(np.concatenate([a, b], 1)[1], got[1])
    assert eng.frames(3) == 7 and eng.frames(0) == 0


def test_persistent_launch_equals_per_phase_launches(E, monkeypatch):
    outs = []
    for dbg in ("0", "1"):
        monkeypatch.setenv("S2S_DEBUG_PHASES", dbg)
        g, cg, w, cw, eng = _engine(E, "bfloat16")
        eng.prefill(0, [5, 6, 7, 8], 2301)
        outs.append(eng.decode_frames([0], 5).cpu().numpy().copy())
        eng.close()
    assert np.array_equal(outs[0], outs[1])


def test_streaming_audio_matches_the_oracle_chunked_decode(E):
    """8-frame chunks behind 25 frames of history (the reference handler's chunk_size, S/TTS/qwen3_tts_handler.py:49):
    the concatenated chunks equal the oracle's chunked_decode of the oracle's codes."""
    g, cg, w, cw, eng = _engine(E, "float16")
    text, F, chunk, left = [11, 12, 13, 14, 15], 20, 8, 25
    codes_ref = R.generate(w, g, text, 2301, F)
    assert len(codes_ref) == F
    wav_ref = C.chunked_decode(cw, cg, codes_ref.T, chunk_size=chunk, left_context=left)
    eng.prefill(0, text, 2301)
    outs, done = [], 0
    while done < F:
        n = min(chunk, F - done)
        eng.decode_frames([0], n)
        outs.append(eng.decode_audio(0, n, left).cpu().numpy())
        done += n
    got = np.concatenate(outs)
    assert got.shape == wav_ref.shape
    assert np.abs(got - wav_ref).max() < 1e-3
