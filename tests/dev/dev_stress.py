"""Developer aid (not a test): run-to-run determinism / race stress of the persistent decode kernels.
Every repetition of the same launch must give bit-identical ids and logits."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from oracle import weights as W
from speech_to_speech_b200 import engine as E
import bench

bad = 0
g = W.WHISPER_GEOMETRIES["small"]
for NB, reps in ((1, 40), (2, 10), (16, 10)):
    eng = E.WhisperEngine(g.to_dict(), max_batch=NB); eng.init_random(1)
    opts = E.WhisperDecodeOptions(prefix=bench.PREFIX, eos_id=-1, max_new_tokens=48, suppress=bench.SUPPRESS, begin_suppress=bench.BEGIN_SUPPRESS)
    pcm = torch.from_numpy(np.stack([W.synthetic_audio(i, 160000) for i in range(NB)])).cuda()
    eng.logmel(pcm, [160000] * NB); eng.encode(NB)
    ref = None
    for r in range(reps):
        ids, lens, lg = eng.decode(NB, opts, return_logits=True)
        cur = (ids.cpu().numpy().copy(), lg.cpu().numpy().copy())
        if ref is None:
            ref = cur
        elif not (np.array_equal(ref[0], cur[0]) and np.array_equal(ref[1], cur[1], equal_nan=True)):
            bad += 1
            print(f"whisper B={NB} rep {r}: ids differ {int((ref[0] != cur[0]).sum())}, logits differ {int((ref[1] != cur[1]).sum())}")
    print(f"whisper B={NB}: {reps} repetitions, mismatches so far {bad}", flush=True)
    eng.close()
gl = W.LLAMA_GEOMETRIES["mini"]
eng = E.LlamaEngine(gl.to_dict(), dtype="bfloat16", max_sessions=4, max_positions=256, max_prefill=64); eng.init_random(2)
prompts = [np.random.default_rng(s).integers(0, gl.vocab, 20 + 7 * s).tolist() for s in range(4)]
ref = None
for r in range(15):
    first = []
    for s, p in enumerate(prompts):
        eng.reset(s)
        nxt, _ = eng.prefill(s, p)
        first.append(int(nxt))
    ids, lens, lg = eng.decode([0, 1, 2, 3], torch.tensor(first, dtype=torch.int32, device="cuda"), 24, return_logits=True)
    cur = (ids.cpu().numpy().copy(), lg.cpu().numpy().copy())
    if ref is None:
        ref = cur
    elif not (np.array_equal(ref[0], cur[0]) and np.array_equal(ref[1], cur[1])):
        bad += 1
        print(f"llama rep {r}: ids differ {int((ref[0] != cur[0]).sum())}, logits differ {int((ref[1] != cur[1]).sum())}")
print(f"llama B=4: 15 repetitions, mismatches total {bad}")
print("STRESS OK" if bad == 0 else f"STRESS FAILED {bad}")
