"""CPU: the numpy oracle (oracle/whisper_ref.py) against the transformers-generated golden vectors
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU tests then compare the CUDA path
with the oracle and the same goldens."""
import os

import numpy as np
import pytest

from oracle import weights as W, whisper_ref as R


@pytest.fixture(scope="module", params=["micro", "tiny"])
def case(request, golden_dir):
    name = request.param
    g = W.WHISPER_GEOMETRIES[name]
    G = np.load(os.path.join(golden_dir, f"whisper_{name}.npz"))
    w = W.make_whisper_weights(g, 0)
    audio = W.synthetic_audio(int(G["audio_seed"]), int(G["n_samples"]))
    mel = R.log_mel_spectrogram(audio, g.n_mels)
    enc, hs = R.encoder_forward(w, g, mel, return_all=True)
    return dict(name=name, g=g, G=G, w=w, mel=mel, enc=enc, hs=hs)


def test_weights_are_fp16_exact():
    w = W.make_whisper_weights(W.WHISPER_GEOMETRIES["micro"], 0)
    for k, v in w.items():
        assert v.dtype == np.float32
        assert np.array_equal(v, v.astype(np.float16).astype(np.float32)), k


def test_logmel_matches_transformers(case):
    G = case["G"]
    np.testing.assert_allclose(case["mel"][:, G["frame_idx"]], G["mel_frames"], atol=2e-5, rtol=0)
    assert abs(case["mel"].max() - float(G["mel_max"])) < 1e-5


def test_encoder_matches_transformers(case):
    G = case["G"]
    np.testing.assert_allclose(case["hs"][0][G["row_idx"]], G["conv_rows"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(case["hs"][1][G["row_idx"]], G["layer0_rows"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(case["enc"][G["row_idx"]], G["enc_rows"], atol=2e-5, rtol=0)


def test_greedy_ids_and_logits_match_transformers(case):
    G, g, w = case["G"], case["g"], case["w"]
    ids, lg = R.greedy_decode(w, g, case["enc"], G["prefix"].tolist(), int(G["max_new"]), int(G["eos"]),
                              G["suppress"].tolist(), G["begin_suppress"].tolist(), return_logits=True)
    assert ids == G["gen_ids"].tolist()  # bit-exact token ids
    tv = np.take_along_axis(lg, G["top_idx"][: len(ids)], 1)
    np.testing.assert_allclose(tv, G["top_val"][: len(ids)], atol=5e-6, rtol=0)
    np.testing.assert_allclose(lg[0][G["col_idx"]], G["step0_cols"], atol=5e-6, rtol=0)


def test_suppress_masks_apply(case):
    G, g, w = case["G"], case["g"], case["w"]
    ids, lg = R.greedy_decode(w, g, case["enc"], G["prefix"].tolist(), 2, int(G["eos"]),
                              G["suppress"].tolist(), G["begin_suppress"].tolist(), return_logits=True)
    assert np.isneginf(lg[0][G["suppress"]]).all() and np.isneginf(lg[0][G["begin_suppress"]]).all()
    assert np.isneginf(lg[1][G["suppress"]]).all() and not np.isneginf(lg[1][G["begin_suppress"][0]])


def test_empty_and_long_audio_features():
    m0 = R.log_mel_spectrogram(np.zeros(0, np.float32), 80)
    assert m0.shape == (80, 3000) and np.allclose(m0, -1.5)
    long = W.synthetic_audio(5, 500000)
    np.testing.assert_array_equal(R.log_mel_spectrogram(long, 80), R.log_mel_spectrogram(long[:480000], 80))


def test_detect_language_is_masked_argmax(case):
    g, w = case["g"], case["w"]
    lang = [5, 17, 33, 250]
    got = R.detect_language(w, g, case["enc"], int(case["G"]["prefix"][0]), lang)
    assert got in lang


def test_detect_language_matches_transformers_detect_language(golden_dir):
    """oracle.detect_language pinned to WhisperGenerationMixin.detect_language (TF generation_whisper.py:1610-1674) on six
    utterances: the same language token, and the same language logits within 1e-4 (tests/golden/make_golden.py langdetect)."""
    import os
    G = np.load(os.path.join(golden_dir, "whisper_langdetect_micro.npz"))
    g = W.WHISPER_GEOMETRIES["micro"]
    w = W.make_whisper_weights(g, 0)
    lang = G["lang_ids"].tolist()
    for seed, n, want, lg in zip(G["audio_seeds"], G["n_samples"], G["detected"], G["lang_logits"]):
        enc = R.encoder_forward(w, g, R.log_mel_spectrogram(W.synthetic_audio(int(seed), int(n)), g.n_mels))
        assert R.detect_language(w, g, enc, int(G["sot"]), lang) == int(want)
        st = R.DecoderState(g, R.cross_kv(w, g, enc))
        logits = R.decoder_step(w, g, st, int(G["sot"]))
        assert np.abs(logits[lang] - lg).max() < 1e-4
