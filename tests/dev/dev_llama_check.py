"""Developer bring-up (not collected): Llama path error levels per dtype."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from oracle import weights as W, llama_ref as R
from speech_to_speech_b200 import engine as E
GOLD = os.path.join(ROOT, "tests", "golden")
for name in ["micro", "mini"]:
    g = W.LLAMA_GEOMETRIES[name]; w = W.make_llama_weights(g, 0)
    G = np.load(os.path.join(GOLD, f"llama_{name}.npz")); gold = G["gen_ids"]
    for dt in ["float16", "bfloat16"]:
        eng = E.LlamaEngine(g.to_dict(), dtype=dt, max_positions=256, max_prefill=128); eng.load_state_dict(w)
        nxt, logits = eng.prefill(0, G["prompt"].tolist(), return_logits=True)
        lg = logits.cpu().numpy()
        e1 = np.abs(lg[-1][G["col_idx"]] - G["prefill_last_cols"])
        first = torch.tensor([int(gold[0])], dtype=torch.int32, device="cuda")
        forced = torch.from_numpy(np.ascontiguousarray(gold[None, 1:])).cuda().int()
        ids, lens, dl = eng.decode([0], first, len(gold) - 1, forced=forced, return_logits=True)
        dl = dl[:, 0].cpu().numpy(); ids = ids[0].cpu().numpy()
        tv = np.take_along_axis(dl, G["top_idx"][1:], 1); e2 = np.abs(tv - G["top_val"][1:])
        agree = (ids == gold[1:]).sum()
        print(f"{name} {dt}: prefill last-row logits err max {e1.max():.4f} mean {e1.mean():.4f} (std {G['prefill_last_cols'].std():.3f}); "
              f"next_id ok {int(nxt[0]) == int(gold[0])}; decode top-8 err max {e2.max():.4f} mean {e2.mean():.4f}; ids agree {agree}/{len(gold)-1}")
