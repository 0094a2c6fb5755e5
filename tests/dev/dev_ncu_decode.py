"""Developer aid (not a test): the Whisper-small decode launch for an ncu capture.  argv[1] = sessions per launch."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from oracle import weights as W
from speech_to_speech_b200 import engine as E
import bench

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 1
g = W.WHISPER_GEOMETRIES["small"]
eng = E.WhisperEngine(g.to_dict(), max_batch=NB); eng.init_random(1)
opts = E.WhisperDecodeOptions(prefix=bench.PREFIX, eos_id=-1, max_new_tokens=128, suppress=bench.SUPPRESS, begin_suppress=bench.BEGIN_SUPPRESS)
pcm = torch.from_numpy(np.stack([W.synthetic_audio(i, 160000) for i in range(NB)])).cuda()
eng.logmel(pcm, [160000] * NB); eng.encode(NB)
for _ in range(3):
    eng.decode(NB, opts)
torch.cuda.synchronize()
