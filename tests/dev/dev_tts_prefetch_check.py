"""Developer aid (not a test): the prefetched TTS chunk requests on the real engine -- 8 concurrent sessions on a shared micro
model (2 lanes x 4), device post-processing, batch sizes reached."""
import os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
from speech_to_speech_b200.tts_model import B200Qwen3TTS
from speech_to_speech_b200.handlers.qwen3_tts_postproc import TTSPostProcessor
sys.setswitchinterval(0.0005)
post = TTSPostProcessor(0)
models = [B200Qwen3TTS.from_random("micro", seed=3, dtype="bfloat16", max_sessions=4, max_positions=256, max_text=64, lane=l, lanes=2) for l in range(2)]
errs, outs = [], {}

def session(i):
    m = models[i % 2]
    try:
        with m.lane_context():
            n, t0 = 0, time.perf_counter()
            for audio, sr, info in m.generate_custom_voice_streaming("hello there", "Aiden", chunk_size=8, max_new_tokens=64):
                n += len(post.from_device(audio.tensor))
            outs[i] = (n, info["frames"], time.perf_counter() - t0)
    except Exception as e:  # noqa: BLE001
        errs.append(f"{i}: {type(e).__name__}: {e}")
for rep in range(2):
    ths = [threading.Thread(target=session, args=(i,)) for i in range(8)]
    [t.start() for t in ths]
    [t.join(60) for t in ths]
print("errors", errs)
print("outs", outs)
for l, m in enumerate(models):
    b = m.batcher
    print(f"lane {l}: launch groups {b.batches_run}, requests {b.items_run}, largest batch {b.largest_batch}")
    m.close()
print("__PREFETCH_OK__" if not errs and len(outs) == 8 else "__PREFETCH_FAIL__")
