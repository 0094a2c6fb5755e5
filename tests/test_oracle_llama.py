"""CPU: the numpy Llama oracle (oracle/llama_ref.py) against the transformers-generated golden vectors."""
import os

import numpy as np
import pytest

from oracle import weights as W, llama_ref as R


@pytest.mark.parametrize("name", ["micro", "mini", "qwen3-micro"])  # qwen3-micro: Qwen3ForCausalLM (qk_norm), oracle only
def test_greedy_ids_logits_and_hidden_match_transformers(golden_dir, name):
    g = W.LLAMA_GEOMETRIES[name]
    w = W.make_llama_weights(g, 0)
    G = np.load(os.path.join(golden_dir, f"llama_{name}.npz"))
    ids, lg = R.greedy_generate(w, g, G["prompt"], int(G["max_new"]), return_logits=True)
    assert ids == G["gen_ids"].tolist()  # bit-exact ids
    np.testing.assert_allclose(np.take_along_axis(lg, G["top_idx"], 1), G["top_val"], atol=1e-4, rtol=0)
    cache = R.KVCache(g)
    logits, hs, xn = R.forward(w, g, G["prompt"], cache, return_hidden=True)
    np.testing.assert_allclose(logits[-1][G["col_idx"]], G["prefill_last_cols"], atol=1e-4, rtol=0)
    np.testing.assert_allclose(logits[len(G["prompt"]) // 2][G["col_idx"]], G["prefill_mid_cols"], atol=1e-4, rtol=0)
    np.testing.assert_allclose(hs[1][-1], G["h1_last"], atol=1e-4, rtol=0)


def test_weights_are_bf16_exact():
    w = W.make_llama_weights(W.LLAMA_GEOMETRIES["micro"], 0)
    for k, v in w.items():
        assert np.array_equal(v, W.round_bf16(v)), k


def test_chunked_prefill_equals_single_pass():
    g = W.LLAMA_GEOMETRIES["micro"]
    w = W.make_llama_weights(g, 0)
    ids = np.random.default_rng(0).integers(0, g.vocab, 20)
    a = R.forward(w, g, ids, R.KVCache(g))
    c = R.KVCache(g)
    b1 = R.forward(w, g, ids[:7], c)
    b2 = R.forward(w, g, ids[7:], c)
    np.testing.assert_allclose(np.concatenate([b1, b2]), a, atol=2e-5, rtol=0)
