"""Generate golden vectors from `transformers` (the package the reference delegates to).

Run in the builder container (transformers 5.5.0, CPU):
    python tests/golden/make_golden.py [whisper|llama|code2wav|all]

The reference holds no golden mel/logits/ids for this path, so the
pin is the upstream model code itself: WhisperFeatureExtractor + WhisperForConditionalGeneration
(what S/STT/whisper_stt_handler.py:71-87,243 calls) and LlamaForCausalLM
(S/LLM/language_model.py:811-817) at seeded random-init weights from oracle/weights.py.
Fixtures are subsampled to stay small; the sampling indices are stored with them.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import weights as W  # noqa: E402

# Decoder prompt / suppress lists used for every random-init Whisper case.  Real checkpoints
# take these from generation_config.json; random-init configs have none, so
# we drive generate() with an explicit 4-token prefix and explicit suppress lists.
WHISPER_CASES = {
    "micro": dict(prefix=[4000, 4001, 4002, 4003], eos=4095, suppress=[1, 2, 7, 8, 9, 10, 14, 25, 4000, 4001],
                  begin_suppress=[220, 4095], max_new=24, audio_seed=0, n_samples=160000),
    "tiny": dict(prefix=[50258, 50259, 50359, 50363], eos=50257,
                 suppress=[1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93,
                           359, 503, 522, 542, 873, 893, 902, 918, 922, 931, 1350, 1853, 1982, 2460, 2627, 3246,
                           3253, 3268, 3536, 3846, 3961, 4183, 4667, 6585, 6647, 7273, 9061, 9383, 10428, 10929,
                           11938, 12033, 12331, 12562, 13793, 14157, 14635, 15265, 15618, 16553, 16604, 18362,
                           18956, 20075, 21675, 22520, 26130, 26161, 26435, 28279, 29464, 31650, 32302, 32470,
                           36865, 42863, 47425, 49870, 50254, 50258, 50358, 50359, 50360, 50361, 50362],
                 begin_suppress=[220, 50257], max_new=32, audio_seed=1, n_samples=160000),
}


def build_hf_whisper(geom: W.WhisperGeometry, weights):
    import torch
    from transformers import WhisperConfig, WhisperForConditionalGeneration

    cfg = WhisperConfig(
        vocab_size=geom.vocab, num_mel_bins=geom.n_mels, d_model=geom.d_model,
        encoder_layers=geom.enc_layers, decoder_layers=geom.dec_layers,
        encoder_attention_heads=geom.heads, decoder_attention_heads=geom.heads,
        encoder_ffn_dim=geom.ffn, decoder_ffn_dim=geom.ffn,
        max_source_positions=geom.max_source_positions, max_target_positions=geom.max_target_positions,
        pad_token_id=0, bos_token_id=0, eos_token_id=0, decoder_start_token_id=0,
        suppress_tokens=None, begin_suppress_tokens=None,
    )
    cfg._attn_implementation = "eager"
    model = WhisperForConditionalGeneration(cfg).eval()
    sd = {k: torch.from_numpy(v.copy()) for k, v in weights.items()}
    sd["proj_out.weight"] = sd["model.decoder.embed_tokens.weight"]
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("proj_out" in m for m in missing), missing
    return model


def whisper_golden(name: str):
    import torch
    from transformers import WhisperFeatureExtractor

    case = WHISPER_CASES[name]
    geom = W.WHISPER_GEOMETRIES[name]
    weights = W.make_whisper_weights(geom, seed=0)
    model = build_hf_whisper(geom, weights)
    audio = W.synthetic_audio(case["audio_seed"], case["n_samples"])

    fe = WhisperFeatureExtractor(feature_size=geom.n_mels)
    feats = fe(audio, sampling_rate=16000, return_tensors="pt").input_features  # [1, n_mels, 3000]
    with torch.no_grad():
        enc = model.get_encoder()(feats, output_hidden_states=True)
        enc_out = enc.last_hidden_state[0].numpy()
        hs = [h[0].numpy() for h in enc.hidden_states]  # conv+pos, then after each layer
        # teacher-free greedy through the public generate() exactly as the handler calls it
        eos = case["eos"]
        model.generation_config.eos_token_id = eos
        model.generation_config.pad_token_id = eos
        model.generation_config.decoder_start_token_id = case["prefix"][0]
        out = model.generate(
            feats,
            decoder_input_ids=torch.tensor([case["prefix"]]),
            max_new_tokens=case["max_new"], num_beams=1, do_sample=False,
            suppress_tokens=case["suppress"], begin_suppress_tokens=case["begin_suppress"],
            return_timestamps=False, output_scores=True, return_dict_in_generate=True,
        )
        seq = out.sequences[0].tolist()
        scores = torch.stack([s[0] for s in out.scores]).numpy()  # processed logits [steps, V]
    n_pref = len(case["prefix"])
    gen_ids = seq[n_pref:] if seq[:n_pref] == case["prefix"] else seq
    gen_ids = gen_ids[: scores.shape[0]]

    mel = feats[0].numpy()
    frame_idx = np.arange(0, 3000, 7)
    row_idx = np.arange(0, 1500, 25)
    top_idx = np.argsort(-scores, axis=1)[:, :8]
    top_val = np.take_along_axis(scores, top_idx, axis=1)
    col_idx = np.arange(0, geom.vocab, 16)
    np.savez_compressed(
        os.path.join(OUT, f"whisper_{name}.npz"),
        mel_frames=mel[:, frame_idx], frame_idx=frame_idx, mel_max=mel.max(), mel_mean=mel.mean(),
        enc_rows=enc_out[row_idx], row_idx=row_idx,
        conv_rows=hs[0][row_idx], layer0_rows=hs[1][row_idx],
        enc_abs_mean=np.abs(enc_out).mean(),
        gen_ids=np.asarray(gen_ids, np.int32),
        top_idx=top_idx.astype(np.int32), top_val=top_val.astype(np.float32),
        step0_cols=scores[0][col_idx], col_idx=col_idx,
        prefix=np.asarray(case["prefix"], np.int32), eos=eos,
        suppress=np.asarray(case["suppress"], np.int32),
        begin_suppress=np.asarray(case["begin_suppress"], np.int32),
        max_new=case["max_new"], audio_seed=case["audio_seed"], n_samples=case["n_samples"],
    )
    print(f"whisper_{name}: ids={gen_ids[:12]}... n={len(gen_ids)} mel[{mel.min():.3f},{mel.max():.3f}]")


def build_hf_llama(geom: W.LlamaGeometry, weights):
    import torch
    from transformers import LlamaConfig, LlamaForCausalLM

    if geom.qk_norm:  # Qwen3 family: same decoder with RMSNorm on q / k heads
        from transformers import Qwen3Config, Qwen3ForCausalLM
        cfg = Qwen3Config(
            vocab_size=geom.vocab, hidden_size=geom.d_model, intermediate_size=geom.ffn, num_hidden_layers=geom.layers,
            num_attention_heads=geom.heads, num_key_value_heads=geom.kv_heads, head_dim=geom.head_dim,
            max_position_embeddings=geom.max_positions, rms_norm_eps=geom.rms_eps, rope_theta=geom.rope_theta,
            tie_word_embeddings=False, attention_bias=False, use_sliding_window=False,
        )
        cfg._attn_implementation = "eager"
        model = Qwen3ForCausalLM(cfg).eval()
        model.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in weights.items()}, strict=True)
        return model
    cfg = LlamaConfig(
        vocab_size=geom.vocab, hidden_size=geom.d_model, intermediate_size=geom.ffn,
        num_hidden_layers=geom.layers, num_attention_heads=geom.heads, num_key_value_heads=geom.kv_heads,
        head_dim=geom.head_dim, max_position_embeddings=geom.max_positions, rms_norm_eps=geom.rms_eps,
        rope_theta=geom.rope_theta, tie_word_embeddings=False, attention_bias=False, mlp_bias=False,
    )
    cfg._attn_implementation = "eager"
    model = LlamaForCausalLM(cfg).eval()
    sd = {k: torch.from_numpy(v.copy()) for k, v in weights.items()}
    model.load_state_dict(sd, strict=True)
    return model


LLAMA_CASES = {
    "micro": dict(prompt_len=24, max_new=16, seed=3),
    "mini": dict(prompt_len=64, max_new=24, seed=4),
    "qwen3-micro": dict(prompt_len=24, max_new=16, seed=5),
}


def llama_prompt(geom: W.LlamaGeometry, n: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, geom.vocab, size=n).astype(np.int32)


def llama_golden(name: str):
    import torch

    case = LLAMA_CASES[name]
    geom = W.LLAMA_GEOMETRIES[name]
    weights = W.make_llama_weights(geom, seed=0)
    model = build_hf_llama(geom, weights)
    prompt = llama_prompt(geom, case["prompt_len"], case["seed"])
    ids = torch.from_numpy(prompt.astype(np.int64))[None]
    with torch.no_grad():
        full = model(ids, output_hidden_states=True)
        prefill_logits = full.logits[0].numpy()  # [P, V]
        h_last = full.hidden_states[-1][0].numpy()
        h1 = full.hidden_states[1][0].numpy()
        out = model.generate(
            ids, max_new_tokens=case["max_new"], min_new_tokens=case["max_new"], do_sample=False,
            output_scores=True, return_dict_in_generate=True, pad_token_id=0,
        )
        gen = out.sequences[0, ids.shape[1]:].numpy().astype(np.int32)
        scores = torch.stack([s[0] for s in out.scores]).numpy()
    top_idx = np.argsort(-scores, axis=1)[:, :8]
    top_val = np.take_along_axis(scores, top_idx, axis=1)
    col_idx = np.arange(0, geom.vocab, 8)
    np.savez_compressed(
        os.path.join(OUT, f"llama_{name}.npz"),
        prompt=prompt, gen_ids=gen, top_idx=top_idx.astype(np.int32), top_val=top_val.astype(np.float32),
        prefill_last_cols=prefill_logits[-1][col_idx], prefill_mid_cols=prefill_logits[len(prompt) // 2][col_idx],
        col_idx=col_idx, h1_last=h1[-1], hN_last=h_last[-1], max_new=case["max_new"],
    )
    print(f"llama_{name}: gen={gen[:10]}...")


def code2wav_golden(name: str = "micro"):
    """Qwen3OmniMoeCode2Wav (the published cousin of the Qwen3-TTS 12 Hz codec decoder; oracle/code2wav_ref.py) at seeded
    weights: codes [Q, T] -> waveform, plus the pre-transformer output."""
    import torch
    from transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe import Qwen3OmniMoeCode2WavConfig
    from transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe import Qwen3OmniMoeCode2Wav
    from oracle import code2wav_ref as C

    g = C.GEOMETRIES[name]
    w = C.make_weights(g, 0)
    cfg = Qwen3OmniMoeCode2WavConfig(codebook_size=g.codebook_size, hidden_size=g.hidden, num_attention_heads=g.heads,
                                     num_key_value_heads=g.kv_heads, intermediate_size=g.inter, num_hidden_layers=g.layers,
                                     num_quantizers=g.quantizers, upsample_rates=g.upsample_rates,
                                     upsampling_ratios=g.upsampling_ratios, decoder_dim=g.decoder_dim,
                                     sliding_window=g.sliding_window, max_position_embeddings=g.max_positions, rms_norm_eps=g.rms_eps)
    m = Qwen3OmniMoeCode2Wav(cfg).eval()
    assert set(m.state_dict().keys()) == set(w.keys())
    m.load_state_dict({k: torch.from_numpy(v) for k, v in w.items()})
    codes = np.random.default_rng(1).integers(0, g.codebook_size, (g.quantizers, 20)).astype(np.int64)
    with torch.no_grad():
        t = torch.from_numpy(codes)[None]
        hidden = m.pre_transformer(inputs_embeds=m.code_embedding(t + m.code_offset).mean(1)).last_hidden_state[0].numpy()
        wav = m(t)[0, 0].numpy()
        chunked = m.chunked_decode(t, chunk_size=8, left_context_size=6)[0, 0].numpy()
    path = os.path.join(OUT, f"code2wav_{name}.npz")
    np.savez_compressed(path, codes=codes, hidden=hidden.astype(np.float32), wav=wav.astype(np.float32),
                        chunked=chunked.astype(np.float32), chunk_size=8, left_context=6, transformers_version=__import__("transformers").__version__)
    print("wrote", path, wav.shape, float(np.abs(wav).max()))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("whisper", "all"):
        for n in WHISPER_CASES:
            whisper_golden(n)
    if what in ("llama", "all"):
        for n in LLAMA_CASES:
            llama_golden(n)
    if what in ("code2wav", "all"):
        code2wav_golden("micro")
