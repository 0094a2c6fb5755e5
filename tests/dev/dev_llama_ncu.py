import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from oracle import weights as W
from speech_to_speech_b200 import engine as E
g = W.LLAMA_GEOMETRIES["llama-3-8b-2l"]
eng = E.LlamaEngine(g.to_dict(), dtype="bfloat16", max_sessions=1, max_positions=256, max_prefill=64); eng.init_random(1)
prompt = np.random.default_rng(0).integers(0, g.vocab, 32).tolist()
nxt, _ = eng.prefill(0, prompt)
for _ in range(2):
    ids, lens = eng.decode([0], nxt, 6)
torch.cuda.synchronize()
