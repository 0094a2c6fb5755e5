"""Developer bring-up script (NOT collected by pytest): runs every CUDA stage against the oracle and prints
per-stage errors instead of stopping at the first failure.  Usage on the GPU box:
    python tests/dev/dev_gpu_check.py [stage ...]      stages: gemm attn logmel encode decode e2e
"""
from __future__ import annotations

import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import weights as W, whisper_ref as R  # noqa: E402
from speech_to_speech_b200 import engine as E  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def stage_gemm():
    torch.manual_seed(0)
    for dt in (torch.float16, torch.bfloat16):
        for (M, N, K) in [(128, 64, 64), (128, 128, 128), (1500, 768, 768), (200, 64, 240), (3000, 2304, 768), (333, 512, 3072)]:
            a = (torch.randn(M, K, device="cuda") * 0.5).to(dt)
            w = (torch.randn(N, K, device="cuda") * 0.05).to(dt)
            bias = torch.randn(N, device="cuda") * 0.1
            ref = a.float() @ w.float().T + bias
            out = E.gemm(a, w, bias, out_dtype=torch.float32)
            torch.cuda.synchronize()
            err = (out - ref).abs().max().item()
            out_h = E.gemm(a, w, bias, act="gelu")
            ref_h = torch.nn.functional.gelu(ref)
            err_h = (out_h.float() - ref_h).abs().max().item()
            print(f"gemm {dt} M{M} N{N} K{K}: max|err| f32-out {err:.3e}  gelu-16bit-out {err_h:.3e}  ref|max| {ref.abs().max().item():.2f}")


def stage_attn():
    torch.manual_seed(1)
    for dt in (torch.float16, torch.bfloat16):
        for (B, T, H, KVH, hd, causal) in [(1, 1500, 6, 6, 64, False), (2, 200, 2, 2, 64, False), (1, 300, 8, 2, 128, True), (1, 64, 4, 4, 64, True)]:
            qkv = (torch.randn(B, T, (H + 2 * KVH) * hd, device="cuda")).to(dt)
            q = qkv[:, :, : H * hd]
            k = qkv[:, :, H * hd : (H + KVH) * hd]
            v = qkv[:, :, (H + KVH) * hd :]
            scale = hd ** -0.5
            o = E.attention(q, k, v, H, KVH, scale, causal)
            torch.cuda.synchronize()
            qf = q.float().view(B, T, H, hd).transpose(1, 2)
            kf = k.float().view(B, T, KVH, hd).transpose(1, 2).repeat_interleave(H // KVH, 1)
            vf = v.float().view(B, T, KVH, hd).transpose(1, 2).repeat_interleave(H // KVH, 1)
            s = qf @ kf.transpose(2, 3) * scale
            if causal:
                s = s + torch.full((T, T), float("-inf"), device="cuda").triu(1)
            ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B, T, H * hd)
            print(f"attn {dt} B{B} T{T} H{H}/{KVH} hd{hd} causal={causal}: max|err| {(o.float() - ref).abs().max().item():.3e}")


def _model(name, max_batch=1, dtype="float16"):
    g = W.WHISPER_GEOMETRIES[name]
    w = W.make_whisper_weights(g, 0)
    eng = E.WhisperEngine(g.to_dict() | {"d_model": g.d_model}, dtype=dtype, max_batch=max_batch)
    eng.load_state_dict(w)
    return g, w, eng


def stage_logmel():
    g, w, eng = _model("micro")
    for seed, n in [(0, 160000), (2, 480000), (3, 12345), (4, 500000)]:
        audio = W.synthetic_audio(seed, n)
        ref = R.log_mel_spectrogram(audio, g.n_mels)
        pcm = torch.from_numpy(audio)[None].cuda().contiguous()
        mel = eng.logmel(pcm, [len(audio)], return_mel=True)
        torch.cuda.synchronize()
        err = np.abs(mel[0].cpu().numpy() - ref)
        print(f"logmel seed{seed} n{n}: max|err| {err.max():.3e} mean {err.mean():.3e}")


def stage_encode(names=("micro", "tiny")):
    for name in names:
        g, w, eng = _model(name)
        G = np.load(os.path.join(GOLD, f"whisper_{name}.npz"))
        audio = W.synthetic_audio(int(G["audio_seed"]), int(G["n_samples"]))
        pcm = torch.from_numpy(audio)[None].cuda().contiguous()
        eng.logmel(pcm, [len(audio)])
        out = eng.encode(1, return_output=True)
        torch.cuda.synchronize()
        o = out[0].cpu().numpy()
        err = np.abs(o[G["row_idx"]] - G["enc_rows"])
        print(f"encode {name}: max|err| {err.max():.3e} mean {err.mean():.3e} (|enc| mean {float(G['enc_abs_mean']):.3f}) finite={np.isfinite(o).all()}")


def stage_decode(names=("micro", "tiny")):
    for name in names:
        g, w, eng = _model(name)
        G = np.load(os.path.join(GOLD, f"whisper_{name}.npz"))
        audio = W.synthetic_audio(int(G["audio_seed"]), int(G["n_samples"]))
        pcm = torch.from_numpy(audio)[None].cuda().contiguous()
        eng.logmel(pcm, [len(audio)])
        eng.encode(1)
        opts = E.WhisperDecodeOptions(prefix=G["prefix"].tolist(), eos_id=int(G["eos"]), max_new_tokens=int(G["max_new"]),
                                      suppress=G["suppress"].tolist(), begin_suppress=G["begin_suppress"].tolist())
        gold = G["gen_ids"]
        forced = torch.from_numpy(np.ascontiguousarray(gold[None, : opts.max_new_tokens])).cuda().int()
        ids, lens, logits = eng.decode(1, opts, forced=forced, return_logits=True)
        torch.cuda.synchronize()
        ids = ids[0].cpu().numpy()
        lg = logits[:, 0].cpu().numpy()
        n = len(gold)
        tv = np.take_along_axis(lg[:n], G["top_idx"][:n], 1)
        lerr = np.abs(tv - G["top_val"][:n]).max()
        margin = G["top_val"][:n, 0] - G["top_val"][:n, 1]
        agree = ids[:n] == gold
        print(f"decode {name} (teacher-forced): logits top-8 max|err| {lerr:.3e}; ids agree {agree.sum()}/{n}; "
              f"min margin where disagree {margin[~agree].min() if (~agree).any() else float('nan'):.3e}; max margin where disagree {margin[~agree].max() if (~agree).any() else float('nan'):.3e}")
        ids2, lens2 = eng.decode(1, opts)
        torch.cuda.synchronize()
        ids2 = ids2[0].cpu().numpy()
        k = 0
        while k < n and ids2[k] == gold[k]:
            k += 1
        print(f"decode {name} (free-running): first {k}/{n} ids equal; len {int(lens2[0])}")


def stage_e2e():
    g, w, eng = _model("tiny")
    G = np.load(os.path.join(GOLD, "whisper_tiny.npz"))
    audio = W.synthetic_audio(int(G["audio_seed"]), int(G["n_samples"]))
    opts = E.WhisperDecodeOptions(prefix=G["prefix"].tolist(), eos_id=int(G["eos"]), max_new_tokens=int(G["max_new"]),
                                  suppress=G["suppress"].tolist(), begin_suppress=G["begin_suppress"].tolist())
    for i in range(3):
        t = time.perf_counter()
        ids = eng.transcribe([audio], opts)
        dt = time.perf_counter() - t
        print(f"e2e tiny transcribe #{i}: {dt * 1e3:.2f} ms, ids[:8]={ids[0][:8]}, gold[:8]={G['gen_ids'][:8].tolist()}")


STAGES = {"gemm": stage_gemm, "attn": stage_attn, "logmel": stage_logmel, "encode": stage_encode, "decode": stage_decode,
          "e2e": stage_e2e}

if __name__ == "__main__":
    which = sys.argv[1:] or list(STAGES)
    print("device:", torch.cuda.get_device_name(0))
    for s in which:
        print(f"===== {s}")
        try:
            STAGES[s]()
        except Exception:
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as e:
                print("cuda context broken:", e)
                break
