"""Developer aid (not a test): cost of the acquire fences in the persistent decode kernels.
A/B over S2S_SYNC_RELAXED (0 = fenced default, 1 = legacy unfenced) x S2S_WHISPER_CLUSTER, Whisper-small 131 steps."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from oracle import weights as W
from speech_to_speech_b200 import engine as E
import bench

g = W.WHISPER_GEOMETRIES["small"]
ref = {}
for NB in (1, 16):
    for cluster in ("1", "0"):
        if NB > 1 and cluster == "1":
            continue
        os.environ["S2S_WHISPER_CLUSTER"] = cluster
        eng = E.WhisperEngine(g.to_dict(), max_batch=NB); eng.init_random(1)
        opts = E.WhisperDecodeOptions(prefix=bench.PREFIX, eos_id=-1, max_new_tokens=128, suppress=bench.SUPPRESS, begin_suppress=bench.BEGIN_SUPPRESS)
        pcm = torch.from_numpy(np.stack([W.synthetic_audio(i, 160000) for i in range(NB)])).cuda()
        eng.logmel(pcm, [160000] * NB); eng.encode(NB)
        for relaxed in ("0", "1"):
            os.environ["S2S_SYNC_RELAXED"] = relaxed
            ids, lens = eng.decode(NB, opts)
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); eng.decode(NB, opts); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            key = NB
            cur = ids.cpu().numpy()
            same = "-" if key not in ref else str(bool(np.array_equal(ref[key], cur)))
            ref.setdefault(key, cur)
            print(f"B={NB} cluster={cluster} relaxed={relaxed}: {min(ts):.2f} ms = {min(ts) / 131 * 1000:.1f} us/step  ids==first:{same}", flush=True)
        eng.close()
os.environ["S2S_SYNC_RELAXED"] = "0"
