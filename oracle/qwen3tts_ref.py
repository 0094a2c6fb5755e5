"""Talker + code predictor oracle of the TTS slot (TEST INFRASTRUCTURE).

The reference's TTS handler calls `faster-qwen3-tts` (`S/TTS/qwen3_tts_handler.py:227, 946-978`:
`model.generate_custom_voice_streaming(text, speaker, language, instruct, chunk_size, max_new_tokens, ...)`), which is
absent from the reference tree, this container and the wheelhouse: **parity with the real Qwen3-TTS is unpinned**.
The nearest published statement of the design is the Qwen3-Omni talker in `transformers`
(TF = site-packages/transformers/models/qwen3_omni_moe/modeling_qwen3_omni_moe.py), which this file restates in numpy:

  text_project        TF:2309-2318  Qwen3OmniMoeTalkerResizeMLP (fc2(silu(fc1 x)), biases)
  build_prompt        TF:3842-3905  _get_talker_assistant_parts: [im_start, assistant, \\n] projected, 4 x tts_pad, tts_bos,
                                    first text token, plus the codec prefix [nothink, think_bos, think_eos, speaker, pad, bos];
                                    the remaining text is fed one token per generated frame (`trailing_text_hidden`), then
                                    tts_eos, then tts_pad
  talker step         TF:3029-3175  Qwen3-style decoder (q/k RMSNorm(head_dim), RoPE; all three M-RoPE sections carry the
                                    same text position, so the rotary table is the ordinary one) + codec_head
  suppress / eos      TF:3954-3973  the last 1024 ids of the codec vocabulary except codec_eos are never predicted
  code predictor      TF:2550-2731  5-layer Qwen3-style decoder: prefill [talker hidden (post final norm), embed(code0)], then
                                    one step per residual codebook, each with its own embedding table and output head
  next talker input   TF:3235-3281  prepare_inputs_for_generation: sum of the frame's 16 code embeddings + the trailing
                                    text embedding (or tts_pad)

Differences from the cousin, all deliberate and stated in DESIGN.md: the talker MLP is DENSE (the real Qwen3-TTS talker is
a dense Qwen3; the Omni talker swaps in a MoE block), and token selection is GREEDY in both models (the cousin samples
with top-k 50 / top-p 0.8 in the code predictor and top-k 50, T 0.9, repetition penalty 1.05 in the talker), because
bit-exact code parity needs a deterministic rule.  Pinned against the transformers classes under exactly these two
substitutions by tests/golden/make_golden.py -> tests/golden/qwen3tts_micro.npz.  float32 throughout.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict, field

import numpy as np

from . import llama_ref as L
from .weights import LlamaGeometry, _rng, round_bf16


@dataclass(frozen=True)
class TTSGeometry:
    talker: LlamaGeometry = field(default_factory=lambda: LlamaGeometry(
        1024, 20, 16, 2, 128, 2048, 3072, rope_theta=1000000.0, rms_eps=1e-6, qk_norm=True))
    predictor: LlamaGeometry = field(default_factory=lambda: LlamaGeometry(
        1024, 5, 16, 8, 128, 3072, 2048, rope_theta=1000000.0, rms_eps=1e-6, qk_norm=True))
    text_vocab: int = 151936          # text-side embedding table (the thinker's in the cousin, the talker's own in Qwen3-TTS)
    text_hidden: int = 2048           # Qwen3OmniMoeTalkerConfig.thinker_hidden_size
    n_groups: int = 16                # codebooks per frame (num_code_groups)
    codec_eos: int = 2150
    codec_nothink: int = 2155
    codec_think_bos: int = 2156
    codec_think_eos: int = 2157
    codec_pad: int = 2148
    codec_bos: int = 2149
    tts_bos: int = 151672             # text-side special tokens (Qwen3OmniMoeConfig.tts_{bos,eos,pad}_token_id)
    tts_eos: int = 151673
    tts_pad: int = 151671
    im_start: int = 151644
    assistant: int = 77091
    newline: int = 198

    def to_dict(self):
        return asdict(self)


GEOMETRIES = {
    "micro": TTSGeometry(
        talker=LlamaGeometry(256, 2, 4, 2, 64, 512, 3072, rope_theta=1000000.0, rms_eps=1e-6, qk_norm=True),
        predictor=LlamaGeometry(256, 2, 4, 2, 64, 512, 2048, rope_theta=1000000.0, rms_eps=1e-6, qk_norm=True),
        text_vocab=512, text_hidden=192, tts_bos=500, tts_eos=501, tts_pad=502, im_start=503, assistant=504, newline=505),
    # Qwen3OmniMoeTalker{Text,CodePredictor}Config defaults with a dense 2048-wide talker MLP (see the module docstring)
    "qwen3-tts-12hz": TTSGeometry(),
}


def suppress_ids(g: TTSGeometry) -> list[int]:
    """TF:3954-3962: special ids of the codec vocabulary that must never be predicted."""
    V = g.talker.vocab
    return [i for i in range(V - 1024, V) if i != g.codec_eos]


# --------------------------------------------------------------------------------------------------- weights
def make_weights(g: TTSGeometry, seed: int = 0) -> dict:
    """Seeded weights under the cousin's state-dict names (`Qwen3OmniMoeTalkerForConditionalGeneration`), plus
    `text_embedding.weight` standing in for the thinker's token embedding.  bf16-representable values."""
    w: dict = {}

    def nrm(name, shape, std, offset=0.0):
        w[name] = round_bf16(offset + _rng(seed, name).standard_normal(shape, dtype=np.float32) * np.float32(std))

    def decoder(prefix: str, gg: LlamaGeometry):
        d = gg.d_model
        for i in range(gg.layers):
            p = f"{prefix}layers.{i}."
            nrm(p + "self_attn.q_proj.weight", (gg.heads * gg.head_dim, d), 2.0 * d ** -0.5)
            nrm(p + "self_attn.k_proj.weight", (gg.kv_heads * gg.head_dim, d), 2.0 * d ** -0.5)
            nrm(p + "self_attn.v_proj.weight", (gg.kv_heads * gg.head_dim, d), 0.8 * d ** -0.5)
            nrm(p + "self_attn.o_proj.weight", (d, gg.heads * gg.head_dim), d ** -0.5)
            nrm(p + "self_attn.q_norm.weight", (gg.head_dim,), 0.1, 1.0)
            nrm(p + "self_attn.k_norm.weight", (gg.head_dim,), 0.1, 1.0)
            nrm(p + "mlp.gate_proj.weight", (gg.ffn, d), d ** -0.5)
            nrm(p + "mlp.up_proj.weight", (gg.ffn, d), d ** -0.5)
            nrm(p + "mlp.down_proj.weight", (d, gg.ffn), 0.7 * gg.ffn ** -0.5)
            nrm(p + "input_layernorm.weight", (d,), 0.1, 1.0)
            nrm(p + "post_attention_layernorm.weight", (d,), 0.1, 1.0)
        nrm(prefix + "norm.weight", (d,), 0.1, 1.0)

    t, c = g.talker, g.predictor
    decoder("model.", t)
    nrm("model.codec_embedding.weight", (t.vocab, t.d_model), 0.5)
    nrm("codec_head.weight", (t.vocab, t.d_model), 0.25)   # wide logits: top-1 margins far above the GPU tolerance
    nrm("text_embedding.weight", (g.text_vocab, g.text_hidden), 1.0)
    nrm("text_projection.linear_fc1.weight", (t.ffn, g.text_hidden), g.text_hidden ** -0.5)
    nrm("text_projection.linear_fc1.bias", (t.ffn,), 0.05)
    nrm("text_projection.linear_fc2.weight", (t.d_model, t.ffn), 0.7 * t.ffn ** -0.5)
    nrm("text_projection.linear_fc2.bias", (t.d_model,), 0.05)
    decoder("code_predictor.model.", c)
    for i in range(g.n_groups - 1):
        nrm(f"code_predictor.model.codec_embedding.{i}.weight", (c.vocab, c.d_model), 0.5)
        nrm(f"code_predictor.lm_head.{i}.weight", (c.vocab, c.d_model), 0.25)
    return w


def talker_view(w: dict) -> dict:
    """The talker's decoder under llama_ref's names."""
    v = {k: a for k, a in w.items() if k.startswith("model.layers.") or k == "model.norm.weight"}
    v["model.embed_tokens.weight"] = w["model.codec_embedding.weight"]
    v["lm_head.weight"] = w["codec_head.weight"]
    return v


def predictor_view(w: dict) -> dict:
    pre = "code_predictor."
    v = {k[len(pre):]: a for k, a in w.items() if k.startswith(pre + "model.layers.") or k == pre + "model.norm.weight"}
    v["model.embed_tokens.weight"] = w["code_predictor.model.codec_embedding.0.weight"]   # unused (inputs are embeddings)
    v["lm_head.weight"] = w["code_predictor.lm_head.0.weight"]                              # unused (heads applied here)
    return v


# --------------------------------------------------------------------------------------------------- arithmetic
def text_project(w: dict, x: np.ndarray) -> np.ndarray:
    """Qwen3OmniMoeTalkerResizeMLP.forward (TF:2317-2318)."""
    h = L.silu(x @ w["text_projection.linear_fc1.weight"].T + w["text_projection.linear_fc1.bias"])
    return (h @ w["text_projection.linear_fc2.weight"].T + w["text_projection.linear_fc2.bias"]).astype(np.float32)


def build_prompt(w: dict, g: TTSGeometry, text_ids, speaker_id: int):
    """TF:3842-3905 for the assistant segment [im_start, assistant, newline] + text_ids (at least one text token).
    -> (inputs_embeds [9, d], trailing [len(text) - 1 + 1, d], tts_pad_embed [d])."""
    ids = np.asarray([g.im_start, g.assistant, g.newline] + list(text_ids), np.int64)
    assert len(text_ids) >= 1
    hid = text_project(w, w["text_embedding.weight"][ids])
    bos, eos, pad = text_project(w, w["text_embedding.weight"][np.asarray([g.tts_bos, g.tts_eos, g.tts_pad])])
    text_part = np.concatenate([hid[:3], np.repeat(pad[None], 4, 0), bos[None], hid[3:4]], 0)
    codec_ids = np.asarray([g.codec_nothink, g.codec_think_bos, g.codec_think_eos, speaker_id, g.codec_pad, g.codec_bos])
    codec_part = np.concatenate([np.zeros((3, g.talker.d_model), np.float32), w["model.codec_embedding.weight"][codec_ids]], 0)
    trailing = np.concatenate([hid[4:], eos[None]], 0)
    return (text_part + codec_part).astype(np.float32), trailing.astype(np.float32), pad.astype(np.float32)


def _decoder(wv: dict, gg: LlamaGeometry, x: np.ndarray, cache: L.KVCache):
    """x [T, d] embeddings appended to the cache -> (logits under wv['lm_head.weight'], post-norm hidden [T, d])."""
    logits, _, xn = L.forward(wv, gg, None, cache, return_hidden=True, inputs_embeds=x)
    return logits, xn


def predict_residual_codes(w: dict, g: TTSGeometry, past_hidden: np.ndarray, code0: int, return_logits: bool = False):
    """TF:3249-3270 with greedy selection: codes 1..n_groups-1 of the frame and the sum of the frame's 16 code embeddings."""
    wv, gg = predictor_view(w), g.predictor
    cache = L.KVCache(gg)
    e0 = w["model.codec_embedding.weight"][code0]
    _, xn = _decoder(wv, gg, np.stack([past_hidden, e0]).astype(np.float32), cache)
    codes, acc, logits_all = [], e0.astype(np.float32).copy(), []
    h = xn[-1]
    for i in range(g.n_groups - 1):
        logits = (h @ w[f"code_predictor.lm_head.{i}.weight"].T).astype(np.float32)
        c = int(np.argmax(logits))
        codes.append(c)
        logits_all.append(logits)
        e = w[f"code_predictor.model.codec_embedding.{i}.weight"][c].astype(np.float32)
        acc = acc + e
        if i + 1 < g.n_groups - 1:
            _, xn = _decoder(wv, gg, e[None], cache)
            h = xn[-1]
    return (codes, acc, np.stack(logits_all)) if return_logits else (codes, acc)


def generate(w: dict, g: TTSGeometry, text_ids, speaker_id: int, max_frames: int, return_logits: bool = False):
    """Greedy talker loop -> codes [F, n_groups] (frames up to, not including, the one whose first code is codec_eos)."""
    wv, gg = talker_view(w), g.talker
    embeds, trailing, pad = build_prompt(w, g, text_ids, speaker_id)
    mask = np.zeros(gg.vocab, bool)
    mask[suppress_ids(g)] = True
    cache = L.KVCache(gg)
    logits, xn = _decoder(wv, gg, embeds, cache)
    frames, t_logits, p_logits = [], [], []
    for f in range(max_frames):
        lg = np.where(mask, -np.inf, logits[-1]).astype(np.float32)
        code0 = int(np.argmax(lg))
        t_logits.append(lg)
        if code0 == g.codec_eos:
            break
        res = predict_residual_codes(w, g, xn[-1], code0, return_logits=return_logits)
        frames.append([code0] + res[0])
        if return_logits:
            p_logits.append(res[2])
        nxt = res[1] + (trailing[f] if f < len(trailing) else pad)
        if f + 1 < max_frames:
            logits, xn = _decoder(wv, gg, nxt[None].astype(np.float32), cache)
    codes = np.asarray(frames, np.int32).reshape(len(frames), g.n_groups)
    if return_logits:
        return codes, np.stack(t_logits), (np.stack(p_logits) if p_logits else np.zeros((0, g.n_groups - 1, g.predictor.vocab), np.float32))
    return codes
