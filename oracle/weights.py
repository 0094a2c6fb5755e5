"""Seeded random-init weights at reference geometries (test infrastructure).

No checkpoints exist offline, so parity runs on seeded
random weights.  Tensor names are the ``transformers`` state-dict names the
reference loads through ``AutoModelForSpeechSeq2Seq`` / ``AutoModelForCausalLM``
(S/STT/whisper_stt_handler.py:71-75, S/LLM/language_model.py:811-817), so the
same dict feeds the HF model (golden generation), this oracle and the CUDA
engine's ``bind_tensor``.

Each tensor draws from its own PCG64 stream keyed by (seed, crc32(name)), so the
values do not depend on creation order.  All values are rounded to the storage
dtype of the CUDA engine (fp16 for Whisper, bf16 for Llama) and returned as
float32, so oracle and engine see bit-identical weights.
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass, asdict

import numpy as np


# --------------------------------------------------------------------------- geometry
@dataclass(frozen=True)
class WhisperGeometry:
    """Mirrors transformers WhisperConfig fields (configuration_whisper.py:127-164)."""

    d_model: int = 384
    heads: int = 6
    enc_layers: int = 4
    dec_layers: int = 4
    ffn: int = 1536
    n_mels: int = 80
    vocab: int = 51865
    max_source_positions: int = 1500
    max_target_positions: int = 448

    def to_dict(self):
        return asdict(self)


WHISPER_GEOMETRIES = {
    # micro: smallest shape the kernels support (head_dim is always 64); for fast CPU tests
    "micro": WhisperGeometry(128, 2, 2, 2, 512, 80, 4096, 1500, 448),
    "tiny": WhisperGeometry(384, 6, 4, 4, 1536, 80, 51865, 1500, 448),
    "small": WhisperGeometry(768, 12, 12, 12, 3072, 80, 51865, 1500, 448),
    "large-v3": WhisperGeometry(1280, 20, 32, 32, 5120, 128, 51866, 1500, 448),
}


@dataclass(frozen=True)
class LlamaGeometry:
    """Mirrors transformers LlamaConfig (modeling_llama.py:225-333)."""

    d_model: int = 4096
    layers: int = 32
    heads: int = 32
    kv_heads: int = 8
    head_dim: int = 128
    ffn: int = 14336
    vocab: int = 128256
    rope_theta: float = 500000.0
    rms_eps: float = 1e-5
    max_positions: int = 8192
    qk_norm: bool = False   # Qwen3: RMSNorm over head_dim on q and k before RoPE (transformers modeling_qwen3.py)

    def to_dict(self):
        return asdict(self)


LLAMA_GEOMETRIES = {
    "micro": LlamaGeometry(256, 2, 2, 1, 128, 512, 2048),
    "mini": LlamaGeometry(1024, 4, 8, 2, 128, 3584, 32064),
    "llama-3-8b-2l": LlamaGeometry(4096, 2, 32, 8, 128, 14336, 128256),
    "llama-3-8b": LlamaGeometry(4096, 32, 32, 8, 128, 14336, 128256),
    # Qwen3 family (the reference's DEFAULT LLM is Qwen/Qwen3-4B-Instruct-2507, language_model_base_arguments.py:6-9):
    # oracle only so far -- the CUDA engine rejects qk_norm until its kernels apply it
    "qwen3-micro": LlamaGeometry(256, 2, 4, 2, 64, 512, 2048, rope_theta=1000000.0, rms_eps=1e-6, qk_norm=True),
}


# --------------------------------------------------------------------------- rounding
def round_fp16(x: np.ndarray) -> np.ndarray:
    return x.astype(np.float16).astype(np.float32)


def round_bf16(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bfloat16, returned as float32."""
    u = np.array(x, dtype=np.float32, order="C", copy=True).view(np.uint32)
    # in-place 32-bit arithmetic (no overflow for finite values: only NaN payloads >= 0xFFFF8000 would wrap);
    # 4x less memory traffic than widening to 64 bits -- the 8B-geometry test tensors have 1.5e9 elements
    lsb = u >> np.uint32(16)
    lsb &= np.uint32(1)
    u += np.uint32(0x7FFF)
    u += lsb
    u &= np.uint32(0xFFFF0000)
    return u.view(np.float32)


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


_BIG = 1 << 26            # tensors above 64 M elements (only the exact-8B test geometry has them) ...
_PERIOD = 16_777_259      # ... repeat a block of this many values; prime, so no two rows / tiles line up


def _normal(seed, name, shape, std, rnd):
    n = int(np.prod(shape))
    if n <= _BIG:
        return rnd(_rng(seed, name).standard_normal(shape, dtype=np.float32) * np.float32(std))
    # generating and rounding 1.5e9 values took minutes of single-threaded numpy per test run; the golden-pinned
    # geometries are far below _BIG and keep their exact streams
    block = rnd(_rng(seed, name).standard_normal(_PERIOD, dtype=np.float32) * np.float32(std))
    return np.resize(block, shape)


def sinusoids(length: int, channels: int, max_timescale: float = 10000.0) -> np.ndarray:
    """Whisper encoder positions (transformers modeling_whisper.py:52-60 `sinusoids`)."""
    log_inc = np.log(max_timescale) / (channels // 2 - 1)
    inv = np.exp(-log_inc * np.arange(channels // 2))
    t = np.arange(length)[:, None] * inv[None, :]
    return np.concatenate([np.sin(t), np.cos(t)], axis=1).astype(np.float32)


# --------------------------------------------------------------------------- whisper
def make_whisper_weights(geom: WhisperGeometry, seed: int = 0, rnd=round_fp16) -> dict[str, np.ndarray]:
    d, f = geom.d_model, geom.ffn
    w: dict[str, np.ndarray] = {}

    def lin(name, out_f, in_f, bias=True, std=0.02):
        w[name + ".weight"] = _normal(seed, name + ".weight", (out_f, in_f), std, rnd)
        if bias:
            w[name + ".bias"] = _normal(seed, name + ".bias", (out_f,), 0.02, rnd)

    def ln(name):
        w[name + ".weight"] = rnd(1.0 + _normal(seed, name + ".weight", (d,), 0.1, lambda a: a))
        w[name + ".bias"] = _normal(seed, name + ".bias", (d,), 0.05, rnd)

    # Scales are chosen (fan-in normalised) so that attention is far from uniform and the layer
    # outputs dominate the token embedding in the residual stream; otherwise the tied
    # output projection just echoes the input token and the test exercises nothing.
    s_qk, s_v, s_o = 1.6 / np.sqrt(d), 0.8 / np.sqrt(d), 1.2 / np.sqrt(d)
    s_fc1, s_fc2 = 1.0 / np.sqrt(d), 0.5 / np.sqrt(f)

    def attn(prefix):
        lin(prefix + ".q_proj", d, d, std=s_qk)
        lin(prefix + ".k_proj", d, d, bias=False, std=s_qk)
        lin(prefix + ".v_proj", d, d, std=s_v)
        lin(prefix + ".out_proj", d, d, std=s_o)

    e = "model.encoder."
    w[e + "conv1.weight"] = _normal(seed, e + "conv1.weight", (d, geom.n_mels, 3), 0.05, rnd)
    w[e + "conv1.bias"] = _normal(seed, e + "conv1.bias", (d,), 0.02, rnd)
    w[e + "conv2.weight"] = _normal(seed, e + "conv2.weight", (d, d, 3), 0.03, rnd)
    w[e + "conv2.bias"] = _normal(seed, e + "conv2.bias", (d,), 0.02, rnd)
    w[e + "embed_positions.weight"] = rnd(sinusoids(geom.max_source_positions, d))
    for i in range(geom.enc_layers):
        p = f"{e}layers.{i}."
        attn(p + "self_attn")
        ln(p + "self_attn_layer_norm")
        lin(p + "fc1", f, d, std=s_fc1)
        lin(p + "fc2", d, f, std=s_fc2)
        ln(p + "final_layer_norm")
    ln(e + "layer_norm")

    dd = "model.decoder."
    w[dd + "embed_tokens.weight"] = _normal(seed, dd + "embed_tokens.weight", (geom.vocab, d), 0.02, rnd)
    w[dd + "embed_positions.weight"] = _normal(
        seed, dd + "embed_positions.weight", (geom.max_target_positions, d), 0.02, rnd
    )
    for i in range(geom.dec_layers):
        p = f"{dd}layers.{i}."
        attn(p + "self_attn")
        ln(p + "self_attn_layer_norm")
        attn(p + "encoder_attn")
        ln(p + "encoder_attn_layer_norm")
        lin(p + "fc1", f, d, std=s_fc1)
        lin(p + "fc2", d, f, std=s_fc2)
        ln(p + "final_layer_norm")
    ln(dd + "layer_norm")
    return w


# --------------------------------------------------------------------------- llama
def make_llama_weights(geom: LlamaGeometry, seed: int = 0, rnd=round_bf16) -> dict[str, np.ndarray]:
    d, f, hd = geom.d_model, geom.ffn, geom.head_dim
    w: dict[str, np.ndarray] = {}

    def lin(name, out_f, in_f, std=0.02):
        w[name + ".weight"] = _normal(seed, name + ".weight", (out_f, in_f), std, rnd)

    def norm(name):
        w[name + ".weight"] = rnd(1.0 + _normal(seed, name + ".weight", (d,), 0.1, lambda a: a))

    w["model.embed_tokens.weight"] = _normal(seed, "model.embed_tokens.weight", (geom.vocab, d), 0.02, rnd)
    for i in range(geom.layers):
        p = f"model.layers.{i}."
        lin(p + "self_attn.q_proj", geom.heads * hd, d, std=2.0 / np.sqrt(d))
        lin(p + "self_attn.k_proj", geom.kv_heads * hd, d, std=2.0 / np.sqrt(d))
        lin(p + "self_attn.v_proj", geom.kv_heads * hd, d, std=0.8 / np.sqrt(d))
        lin(p + "self_attn.o_proj", d, geom.heads * hd, std=1.0 / np.sqrt(d))
        if geom.qk_norm:
            w[p + "self_attn.q_norm.weight"] = rnd(1.0 + _normal(seed, p + "self_attn.q_norm.weight", (hd,), 0.1, lambda a: a))
            w[p + "self_attn.k_norm.weight"] = rnd(1.0 + _normal(seed, p + "self_attn.k_norm.weight", (hd,), 0.1, lambda a: a))
        lin(p + "mlp.gate_proj", f, d, std=1.0 / np.sqrt(d))
        lin(p + "mlp.up_proj", f, d, std=1.0 / np.sqrt(d))
        lin(p + "mlp.down_proj", d, f, std=0.7 / np.sqrt(f))
        norm(p + "input_layernorm")
        norm(p + "post_attention_layernorm")
    norm("model.norm")
    lin("lm_head", geom.vocab, d, std=0.02)
    return w


# --------------------------------------------------------------------------- synthetic audio
def synthetic_audio(seed: int, n_samples: int = 160000) -> np.ndarray:
    """sum of 5 sinusoids 100-3400 Hz with random phase + 0.01 noise, 16 kHz mono f32."""
    rng = np.random.default_rng(1234 + seed)
    t = np.arange(n_samples, dtype=np.float64) / 16000.0
    x = np.zeros(n_samples, dtype=np.float64)
    for _ in range(5):
        fr = rng.uniform(100.0, 3400.0)
        ph = rng.uniform(0, 2 * np.pi)
        x += 0.08 * np.sin(2 * np.pi * fr * t + ph) * (0.5 + 0.5 * np.sin(2 * np.pi * rng.uniform(0.5, 3.0) * t))
    x += 0.01 * rng.standard_normal(n_samples)
    return np.clip(x, -1.0, 1.0).astype(np.float32)
