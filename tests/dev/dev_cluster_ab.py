"""Developer aid (not a test): A/B of the 8-phase decode kernel and the cluster kernel (S2S_WHISPER_CLUSTER=0/1)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from oracle import weights as W
from speech_to_speech_b200 import engine as E
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "small"
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = W.WHISPER_GEOMETRIES[name]
res = {}
for mode in ("0", "1"):
    os.environ["S2S_WHISPER_CLUSTER"] = mode
    eng = E.WhisperEngine(g.to_dict(), max_batch=NB); eng.init_random(1)
    opts = E.WhisperDecodeOptions(prefix=bench.PREFIX, eos_id=-1, max_new_tokens=128, suppress=bench.SUPPRESS, begin_suppress=bench.BEGIN_SUPPRESS)
    pcm = torch.from_numpy(np.stack([W.synthetic_audio(i, 160000) for i in range(NB)])).cuda()
    eng.logmel(pcm, [160000] * NB); eng.encode(NB)
    ids, lens = eng.decode(NB, opts)
    forced = ids.clone()
    o16 = E.WhisperDecodeOptions(prefix=bench.PREFIX, eos_id=-1, max_new_tokens=16, suppress=bench.SUPPRESS, begin_suppress=bench.BEGIN_SUPPRESS)
    _, _, lg = eng.decode(NB, o16, forced=forced[:, :16].contiguous(), return_logits=True)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.decode(NB, opts); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    res[mode] = (ids.cpu().numpy(), lg.cpu().numpy(), min(ts))
    print(f"cluster={mode}: {min(ts):.2f} ms / 131 steps = {min(ts) / 131 * 1000:.1f} us/step", flush=True)
    eng.close()
i0, l0, _ = res["0"]; i1, l1, _ = res["1"]
fin = np.isfinite(l0)
print("ids equal:", int((i0 == i1).sum()), "/", i0.size, " max |dlogit| (first 16 steps):", float(np.abs(l0[fin] - l1[fin]).max()))
