"""Codec-token -> waveform decoder oracle (TEST INFRASTRUCTURE; groundwork for SURVEY.md §8 row a17).

The reference's TTS slot calls `faster-qwen3-tts` (`S/TTS/qwen3_tts_handler.py:227, 930-1001`), which is absent from the
reference tree, this container and the wheelhouse, so the arithmetic of the real Qwen3-TTS-Tokenizer-12Hz decoder cannot
be pinned here.  The nearest published statement of that design is `Qwen3OmniMoeCode2Wav` in `transformers`
(`TF/models/qwen3_omni_moe/modeling_qwen3_omni_moe.py:3283-3790`): 16 codebooks x 2048 entries -> mean of the code
embeddings -> 8 sliding-window (72) RoPE transformer layers with layer scale -> 2 x (transposed conv x2 + ConvNeXt
block) -> HiFi-GAN-style decoder (causal conv 7, four [SnakeBeta, transposed conv x(8,5,4,3), three dilated residual
units] blocks, SnakeBeta, causal conv 7) -> clamp: 1920 samples per 12.5 Hz frame at 24 kHz, the figures the
handler is built around (`qwen3_tts_handler.py:46-56`).  This file restates THAT module in numpy, function by function,
and is pinned against it by `tests/golden/code2wav_micro.npz` (tests/test_oracle_code2wav.py).
**Parity with the real upstream is unpinned** (its equivalence to this module is an inference); no CUDA
path is built on it yet.  fp32 throughout; layouts follow the torch module ([C, T] for the convolutional part).
"""
from __future__ import annotations

from dataclasses import dataclass, asdict

import numpy as np
from scipy.special import erf


@dataclass(frozen=True)
class Code2WavGeometry:
    codebook_size: int = 2048
    hidden: int = 1024
    heads: int = 16
    kv_heads: int = 16
    inter: int = 3072
    layers: int = 8
    quantizers: int = 16
    upsample_rates: tuple = (8, 5, 4, 3)
    upsampling_ratios: tuple = (2, 2)
    decoder_dim: int = 1536
    sliding_window: int = 72
    rope_theta: float = 10000.0
    rms_eps: float = 1e-5
    max_positions: int = 8000

    def to_dict(self):
        return asdict(self)

    @property
    def total_upsample(self) -> int:
        return int(np.prod(self.upsample_rates + self.upsampling_ratios))


GEOMETRIES = {
    "micro": Code2WavGeometry(codebook_size=64, hidden=64, heads=4, kv_heads=4, inter=96, layers=2, quantizers=4,
                              upsample_rates=(4, 3), upsampling_ratios=(2,), decoder_dim=48, sliding_window=6, max_positions=256),
    "qwen3-12hz": Code2WavGeometry(),  # Qwen3OmniMoeCode2WavConfig defaults
}


# --------------------------------------------------------------------------------------------------- building blocks
# Operand rounding of the dense contractions (linears, convolutions with groups == 1, transposed convolutions).  None: fp32
# throughout (the reference's arithmetic).  np.float16: both operands rounded to fp16, products summed in fp32 -- an emulation
# of the CUDA library's default tensor-core mode (codec precision 1), used by the tests to tell operand-rounding error
# (expected, bounded by this emulation) from a kernel bug.  Set with `operand_rounding(np.float16)`.
_OPERAND = None


class operand_rounding:
    def __init__(self, dtype): self.dtype = dtype
    def __enter__(self):
        global _OPERAND
        self.prev, _OPERAND = _OPERAND, self.dtype
    def __exit__(self, *exc):
        global _OPERAND
        _OPERAND = self.prev


def _q(a: np.ndarray) -> np.ndarray:
    return a if _OPERAND is None else a.astype(_OPERAND).astype(np.float32)


def _lin(h: np.ndarray, wt: np.ndarray) -> np.ndarray:
    """h [T, K] @ wt[N, K].T"""
    return _q(h) @ _q(wt).T


def causal_conv1d(x: np.ndarray, w: np.ndarray, b: np.ndarray, dilation: int = 1, stride: int = 1, groups: int = 1) -> np.ndarray:
    """Qwen3OmniMoeCausalConvNet.forward (:3283-3316): left pad (k_eff - stride), right pad up to a whole frame, conv1d.
    x [C_in, T], w [C_out, C_in / groups, k] -> [C_out, T_out]."""
    c_out, c_in_g, k = w.shape
    k_eff = (k - 1) * dilation + 1
    pad = k_eff - stride
    T = x.shape[-1]
    n_frames = (T - k_eff + pad) / stride + 1
    ideal = (int(np.ceil(n_frames)) - 1) * stride + (k_eff - pad)
    xp = np.pad(x, ((0, 0), (pad, ideal - T)))
    t_out = (xp.shape[-1] - k_eff) // stride + 1
    # windows [C_in, t_out, k] with dilation
    idx = (np.arange(t_out) * stride)[:, None] + (np.arange(k) * dilation)[None, :]
    win = xp[:, idx]                                   # [C_in, t_out, k]
    og = c_out // groups
    out = np.empty((c_out, t_out), np.float32)
    for g in range(groups):
        wg, xg = w[g * og:(g + 1) * og], win[g * c_in_g:(g + 1) * c_in_g]
        if groups == 1:
            wg, xg = _q(wg), _q(xg)                      # depthwise convolutions stay fp32 on the device as well
        out[g * og:(g + 1) * og] = np.einsum("ock,ctk->ot", wg, xg, optimize=True)
    return (out + b[:, None]).astype(np.float32)


def causal_trans_conv1d(x: np.ndarray, w: np.ndarray, b: np.ndarray, stride: int) -> np.ndarray:
    """Qwen3OmniMoeCausalTransConvNet.forward (:3319-3331): ConvTranspose1d then trim ceil(k - stride) from both ends.
    x [C_in, T], w [C_in, C_out, k] -> [C_out, T * stride - (k - stride)]."""
    c_in, c_out, k = w.shape
    T = x.shape[-1]
    full = np.zeros((c_out, (T - 1) * stride + k), np.float32)
    contrib = np.einsum("ct,cok->otk", _q(x), _q(w), optimize=True)   # [C_out, T, k]
    for j in range(k):
        full[:, j:j + (T - 1) * stride + 1:stride] += contrib[:, :, j]
    full += b[:, None]
    pad = int(np.ceil(k - stride))
    return full[:, pad:full.shape[-1] - pad].astype(np.float32)


def snake_beta(x: np.ndarray, alpha: np.ndarray, beta: np.ndarray) -> np.ndarray:
    """SnakeBeta.forward: x + 1 / (exp(beta) + 1e-9) * sin(x * exp(alpha))^2, per channel ([C, T])."""
    a = np.exp(alpha.astype(np.float32))[:, None]
    bb = np.exp(beta.astype(np.float32))[:, None]
    return (x + (np.float32(1.0) / (bb + np.float32(1e-9))) * np.sin(x * a) ** 2).astype(np.float32)


def layer_norm(x: np.ndarray, w: np.ndarray, b: np.ndarray, eps: float) -> np.ndarray:
    m = x.mean(-1, keepdims=True)
    v = ((x - m) ** 2).mean(-1, keepdims=True)
    return ((x - m) / np.sqrt(v + eps) * w + b).astype(np.float32)


def rms_norm(x: np.ndarray, w: np.ndarray, eps: float) -> np.ndarray:
    return (w * (x * (1.0 / np.sqrt((x.astype(np.float32) ** 2).mean(-1, keepdims=True) + eps)))).astype(np.float32)


def gelu(x: np.ndarray) -> np.ndarray:
    return (0.5 * x * (1.0 + erf(x / np.sqrt(2.0)))).astype(np.float32)


def silu(x: np.ndarray) -> np.ndarray:
    return (x / (1.0 + np.exp(-x))).astype(np.float32)


def convnext_block(x: np.ndarray, w: dict, p: str) -> np.ndarray:
    """Qwen3OmniMoeConvNeXtBlock.forward (:3334-3367): depthwise causal conv 7 -> LayerNorm(1e-6) -> 4x MLP (GELU) -> gamma."""
    dim = x.shape[0]
    h = causal_conv1d(x, w[p + "dwconv.conv.weight"], w[p + "dwconv.conv.bias"], groups=dim)
    h = layer_norm(h.T, w[p + "norm.weight"], w[p + "norm.bias"], 1e-6)
    h = gelu(_lin(h, w[p + "pwconv1.weight"]) + w[p + "pwconv1.bias"])
    h = _lin(h, w[p + "pwconv2.weight"]) + w[p + "pwconv2.bias"]
    return (x + (w[p + "gamma"] * h).T).astype(np.float32)


def rope_tables(g: Code2WavGeometry, T: int):
    hd = g.hidden // g.heads
    inv = 1.0 / (g.rope_theta ** (np.arange(0, hd, 2, dtype=np.float32) / hd))
    f = np.arange(T, dtype=np.float32)[:, None] * inv[None, :]
    emb = np.concatenate([f, f], axis=-1)
    return np.cos(emb).astype(np.float32), np.sin(emb).astype(np.float32)


def _rotate_half(x):
    h = x.shape[-1] // 2
    return np.concatenate([-x[..., h:], x[..., :h]], axis=-1)


def transformer_layer(x: np.ndarray, w: dict, p: str, g: Code2WavGeometry, cos, sin) -> np.ndarray:
    """Qwen3OmniMoeCode2WavTransformerLayer.forward (:3494-3554): pre-norm sliding-window attention and SwiGLU MLP,
    each scaled by a LayerScale vector before the residual add."""
    T = x.shape[0]
    hd = g.hidden // g.heads
    grp = g.heads // g.kv_heads
    h = rms_norm(x, w[p + "input_layernorm.weight"], g.rms_eps)
    q = _lin(h, w[p + "self_attn.q_proj.weight"]).reshape(T, g.heads, hd).transpose(1, 0, 2)
    k = _lin(h, w[p + "self_attn.k_proj.weight"]).reshape(T, g.kv_heads, hd).transpose(1, 0, 2)
    v = _lin(h, w[p + "self_attn.v_proj.weight"]).reshape(T, g.kv_heads, hd).transpose(1, 0, 2)
    q = q * cos[None] + _rotate_half(q) * sin[None]
    k = k * cos[None] + _rotate_half(k) * sin[None]
    k = np.repeat(k, grp, axis=0)
    v = np.repeat(v, grp, axis=0)
    s = np.einsum("htd,hsd->hts", q, k, optimize=True) * np.float32(hd ** -0.5)
    i, j = np.arange(T)[:, None], np.arange(T)[None, :]
    visible = (j <= i) & (i - j < g.sliding_window)      # create_sliding_window_causal_mask
    s = np.where(visible[None], s, -np.inf)
    s = s - s.max(-1, keepdims=True)
    pr = np.exp(s)
    pr = pr / pr.sum(-1, keepdims=True)
    a = np.einsum("hts,hsd->thd", pr, v, optimize=True).reshape(T, g.heads * hd)
    x = x + w[p + "self_attn_layer_scale.scale"] * _lin(a, w[p + "self_attn.o_proj.weight"])
    h = rms_norm(x, w[p + "post_attention_layernorm.weight"], g.rms_eps)
    m = _lin(silu(_lin(h, w[p + "mlp.gate_proj.weight"])) * _lin(h, w[p + "mlp.up_proj.weight"]), w[p + "mlp.down_proj.weight"])
    return (x + w[p + "mlp_layer_scale.scale"] * m).astype(np.float32)


def residual_unit(x: np.ndarray, w: dict, p: str, dilation: int) -> np.ndarray:
    """Qwen3OmniMoeCode2WavDecoderResidualUnit.forward (:3686-3702)."""
    h = snake_beta(x, w[p + "act1.alpha"], w[p + "act1.beta"])
    h = causal_conv1d(h, w[p + "conv1.conv.weight"], w[p + "conv1.conv.bias"], dilation=dilation)
    h = snake_beta(h, w[p + "act2.alpha"], w[p + "act2.beta"])
    h = causal_conv1d(h, w[p + "conv2.conv.weight"], w[p + "conv2.conv.bias"])
    return (h + x).astype(np.float32)


# --------------------------------------------------------------------------------------------------- the module
def pre_transformer(w: dict, g: Code2WavGeometry, codes: np.ndarray) -> np.ndarray:
    """codes [Q, T] int -> hidden [T, H]: mean code embedding + transformer + final RMSNorm (:3730-3777, :3557-3642)."""
    Q, T = codes.shape
    assert Q == g.quantizers
    idx = codes + (np.arange(Q) * g.codebook_size)[:, None]
    x = w["code_embedding.weight"][idx].mean(0).astype(np.float32)           # [T, H]
    cos, sin = rope_tables(g, T)
    for l in range(g.layers):
        x = transformer_layer(x, w, f"pre_transformer.layers.{l}.", g, cos, sin)
    return rms_norm(x, w["pre_transformer.norm.weight"], g.rms_eps)


def code2wav_forward(w: dict, g: Code2WavGeometry, codes: np.ndarray, return_hidden: bool = False):
    """Qwen3OmniMoeCode2Wav.forward for one sequence: codes [Q, T] -> wav [T * total_upsample - trims] in [-1, 1]."""
    hidden = pre_transformer(w, g, codes)
    x = hidden.T                                                             # [H, T]
    for i, factor in enumerate(g.upsampling_ratios):
        x = causal_trans_conv1d(x, w[f"upsample.{i}.0.conv.weight"], w[f"upsample.{i}.0.conv.bias"], factor)
        x = convnext_block(x, w, f"upsample.{i}.1.")
    x = causal_conv1d(x, w["decoder.0.conv.weight"], w["decoder.0.conv.bias"])
    for i, rate in enumerate(g.upsample_rates):
        p = f"decoder.{i + 1}.block."
        x = snake_beta(x, w[p + "0.alpha"], w[p + "0.beta"])
        x = causal_trans_conv1d(x, w[p + "1.conv.weight"], w[p + "1.conv.bias"], rate)
        for u, dil in enumerate((1, 3, 9)):
            x = residual_unit(x, w, p + f"{u + 2}.", dil)
    n = len(g.upsample_rates)
    x = snake_beta(x, w[f"decoder.{n + 1}.alpha"], w[f"decoder.{n + 1}.beta"])
    x = causal_conv1d(x, w[f"decoder.{n + 2}.conv.weight"], w[f"decoder.{n + 2}.conv.bias"])
    wav = np.clip(x[0], -1.0, 1.0).astype(np.float32)
    return (wav, hidden) if return_hidden else wav


def chunked_decode(w: dict, g: Code2WavGeometry, codes: np.ndarray, chunk_size: int = 300, left_context: int = 25) -> np.ndarray:
    """Qwen3OmniMoeCode2Wav.chunked_decode (:3779-3790): decode `chunk_size` frames at a time with `left_context` frames
    of history, drop the history's samples -- the streaming form a TTS handler consumes chunk by chunk."""
    T = codes.shape[-1]
    outs, start = [], 0
    while start < T:
        end = min(start + chunk_size, T)
        ctx = left_context if start - left_context > 0 else start
        wav = code2wav_forward(w, g, codes[:, start - ctx:end])
        outs.append(wav[ctx * g.total_upsample:])
        start = end
    return np.concatenate(outs)


# --------------------------------------------------------------------------------------------------- seeded weights
def make_weights(g: Code2WavGeometry, seed: int = 0) -> dict:
    """Seeded fp32 weights under the torch module's state-dict names (same per-tensor PCG64 streams as oracle/weights.py).
    Scales are chosen so that every block matters (the module's own init makes gamma 1e-6 and layer scales 0.01)."""
    from .weights import _rng

    w: dict = {}

    def nrm(name, shape, std, offset=0.0):
        w[name] = (offset + _rng(seed, name).standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)

    H = g.hidden
    hd = H // g.heads
    for l in range(g.layers):
        p = f"pre_transformer.layers.{l}."
        nrm(p + "self_attn.q_proj.weight", (g.heads * hd, H), H ** -0.5)
        nrm(p + "self_attn.k_proj.weight", (g.kv_heads * hd, H), H ** -0.5)
        nrm(p + "self_attn.v_proj.weight", (g.kv_heads * hd, H), H ** -0.5)
        nrm(p + "self_attn.o_proj.weight", (H, g.heads * hd), H ** -0.5)
        nrm(p + "mlp.gate_proj.weight", (g.inter, H), H ** -0.5)
        nrm(p + "mlp.up_proj.weight", (g.inter, H), H ** -0.5)
        nrm(p + "mlp.down_proj.weight", (H, g.inter), g.inter ** -0.5)
        nrm(p + "input_layernorm.weight", (H,), 0.1, 1.0)
        nrm(p + "post_attention_layernorm.weight", (H,), 0.1, 1.0)
        nrm(p + "self_attn_layer_scale.scale", (H,), 0.1, 0.5)
        nrm(p + "mlp_layer_scale.scale", (H,), 0.1, 0.5)
    nrm("pre_transformer.norm.weight", (H,), 0.1, 1.0)
    nrm("code_embedding.weight", (g.codebook_size * g.quantizers, H), 1.0)
    for i, factor in enumerate(g.upsampling_ratios):
        p = f"upsample.{i}."
        nrm(p + "0.conv.weight", (H, H, factor), (H) ** -0.5)
        nrm(p + "0.conv.bias", (H,), 0.02)
        nrm(p + "1.gamma", (H,), 0.05, 0.3)
        nrm(p + "1.dwconv.conv.weight", (H, 1, 7), 7 ** -0.5)
        nrm(p + "1.dwconv.conv.bias", (H,), 0.02)
        nrm(p + "1.norm.weight", (H,), 0.1, 1.0)
        nrm(p + "1.norm.bias", (H,), 0.05)
        nrm(p + "1.pwconv1.weight", (4 * H, H), H ** -0.5)
        nrm(p + "1.pwconv1.bias", (4 * H,), 0.02)
        nrm(p + "1.pwconv2.weight", (H, 4 * H), (4 * H) ** -0.5)
        nrm(p + "1.pwconv2.bias", (H,), 0.02)
    D = g.decoder_dim
    nrm("decoder.0.conv.weight", (D, H, 7), (7 * H) ** -0.5)
    nrm("decoder.0.conv.bias", (D,), 0.02)
    for i, rate in enumerate(g.upsample_rates):
        cin, cout = D // 2 ** i, D // 2 ** (i + 1)
        p = f"decoder.{i + 1}.block."
        nrm(p + "0.alpha", (cin,), 0.2)
        nrm(p + "0.beta", (cin,), 0.2)
        nrm(p + "1.conv.weight", (cin, cout, 2 * rate), (2 * cin) ** -0.5)
        nrm(p + "1.conv.bias", (cout,), 0.02)
        for u in range(3):
            q = p + f"{u + 2}."
            nrm(q + "act1.alpha", (cout,), 0.2)
            nrm(q + "act1.beta", (cout,), 0.2)
            nrm(q + "conv1.conv.weight", (cout, cout, 7), (7 * cout) ** -0.5)
            nrm(q + "conv1.conv.bias", (cout,), 0.02)
            nrm(q + "act2.alpha", (cout,), 0.2)
            nrm(q + "act2.beta", (cout,), 0.2)
            nrm(q + "conv2.conv.weight", (cout, cout, 1), cout ** -0.5)
            nrm(q + "conv2.conv.bias", (cout,), 0.02)
    n = len(g.upsample_rates)
    cl = D // 2 ** n
    nrm(f"decoder.{n + 1}.alpha", (cl,), 0.2)
    nrm(f"decoder.{n + 1}.beta", (cl,), 0.2)
    nrm(f"decoder.{n + 2}.conv.weight", (1, cl, 7), 0.02 * (7 * cl) ** -0.5)  # keeps most of the waveform inside the clamp
    nrm(f"decoder.{n + 2}.conv.bias", (1,), 0.02)
    return w
