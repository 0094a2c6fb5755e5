"""Developer aid (not a test): per-phase timeline of the persistent decode kernel from %globaltimer stamps."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from oracle import weights as W
from speech_to_speech_b200 import engine as E
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "small"
g = W.WHISPER_GEOMETRIES[name]
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = E.WhisperEngine(g.to_dict(), max_batch=NB); eng.init_random(1)
opts = E.WhisperDecodeOptions(prefix=bench.PREFIX, eos_id=-1, max_new_tokens=16, suppress=bench.SUPPRESS, begin_suppress=bench.BEGIN_SUPPRESS)
pcm = torch.from_numpy(np.stack([W.synthetic_audio(i, 160000) for i in range(NB)])).cuda()
eng.logmel(pcm, [160000] * NB); eng.encode(NB)
for _ in range(2): eng.decode(NB, opts)
cap = 4000
tr = torch.zeros((2, cap, 6), dtype=torch.int64, device="cuda")
eng.set_trace(tr); eng.decode(NB, opts); torch.cuda.synchronize(); eng.set_trace(None)
t = tr.cpu().numpy().astype(np.int64)
L = g.dec_layers; nph = 8 * L + 2
names = ["qkv", "self_attn", "self_out", "cross_q", "cross_attn", "cross_out", "fc1", "fc2"]
# steps 0..2 have no logits phase (prefix): phases per step = nph-1; from step 3 on nph
idx = 3 * (nph - 1)  # first full step
for cta in (0, 1):
    body = t[cta, :, 4] - t[cta, :, 0]; bar = t[cta, :, 5] - t[cta, :, 4]; stg = np.where(t[cta, :, 1] > 0, t[cta, :, 1] - t[cta, :, 0], 0)
    print(f"--- CTA {'0' if cta == 0 else 'last'}: step 3.. per-phase mean over layers (ns): body / barrier-wait")
    seg = slice(idx, idx + 4 * nph)
    for k in range(8):
        sel = [idx + s * nph + l * 8 + k for s in range(4) for l in range(L)]
        stg = np.where(t[cta, sel, 1] > 0, t[cta, sel, 1] - t[cta, sel, 0], 0).mean()
        arr = (t[cta, sel, 2] - t[cta, sel, 4]).mean(); prep = (t[cta, sel, 3] - t[cta, sel, 2]).mean(); poll = (t[cta, sel, 5] - t[cta, sel, 3]).mean()
        print(f"  {names[k]:10s} body {body[sel].mean():8.0f}  barrier {bar[sel].mean():8.0f}   [inputs staged after {stg:5.0f} | arrive {arr:5.0f} | prepare next {prep:5.0f} | poll {poll:5.0f}]")
    sel = [idx + s * nph + 8 * L for s in range(4)]
    print(f"  {'logits':10s} body {body[sel].mean():8.0f}  barrier {bar[sel].mean():8.0f}")
    sel = [idx + s * nph + 8 * L + 1 for s in range(4)]
    print(f"  {'select':10s} body {body[sel].mean():8.0f}  barrier {bar[sel].mean():8.0f}")
    tot = t[cta, idx + 4 * nph - 1, 5] - t[cta, idx, 0]
    print(f"  4 steps: {tot / 4e3:.1f} us/step")
print("skew CTA0 vs last at phase begin (ns):", (t[1, idx:idx + 16, 0] - t[0, idx:idx + 16, 0]).tolist())
