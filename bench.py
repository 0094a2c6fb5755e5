#!/usr/bin/env python
"""bench.py -- the headline measurement (contract: task brief "Measurement").

Workload (BASELINE.json configs[1]): Whisper-small, one synthetic 10 s / 16 kHz utterance per step, the whole STT
device path of WhisperSTTHandler.process: log-mel -> encoder (30 s padded window, as the reference) -> greedy
decode of exactly 128 new tokens (4-token forced prompt, suppress lists, no early EOS on either arm).
Weights: seeded random-init at the exact geometry (no checkpoints offline); timing is value independent.

metric  = concurrent real-time sessions = (utterances / s) x 10 s of audio per utterance, whole job over N GPUs
value   : inputs resident in HBM, CUDA-event timed          e2e: host PCM -> ids on host through the C ABI
roofline: the persistent decode kernel (dominant, HBM-bound): algorithmic bytes / measured launch time
cpu_baseline / --impl reference: transformers fp32 on the host cores (the calls the reference handler makes).

    python bench.py --gpus 1 --steps 20 --warmup 3
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 2 --warmup 1
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "small"
AUDIO_S = 10.0
N_SAMPLES = 160000
MAX_NEW = 128
PREFIX = [50258, 50259, 50359, 50363]
SUPPRESS = [1, 2, 7, 8, 9, 10, 14, 25, 26, 27, 28, 29, 31, 58, 59, 60, 61, 62, 63, 90, 91, 92, 93, 359, 503, 522, 542,
            873, 893, 902, 918, 922, 931, 1350, 1853, 1982, 2460, 2627, 3246, 3253, 3268, 3536, 3846, 3961, 4183, 4667,
            6585, 6647, 7273, 9061, 9383, 10428, 10929, 11938, 12033, 12331, 12562, 13793, 14157, 14635, 15265, 15618,
            16553, 16604, 18362, 18956, 20075, 21675, 22520, 26130, 26161, 26435, 28279, 29464, 31650, 32302, 32470,
            36865, 42863, 47425, 49870, 50254, 50258, 50358, 50359, 50360, 50361, 50362]
BEGIN_SUPPRESS = [220, 50257]
METRIC = "concurrent real-time sessions (10 s utterances transcribed per 10 s; Whisper-small STT turn)"


def decode_bytes_per_launch(g, n_prefix, max_new) -> dict:
    """Algorithmic HBM bytes of one persistent-decode launch (DESIGN.md 'whisper_decode_kernel')."""
    d, L, f, V, T = g.d_model, g.dec_layers, g.ffn, g.vocab, g.max_source_positions
    w_layer = (3 * d * d + d * d + d * d + d * d + f * d + d * f) * 2          # 16-bit weights streamed per token
    cross_kv = L * T * 2 * d * 2                                                  # per utterance per token
    logits = V * d * 2
    steps = n_prefix - 1 + max_new
    self_kv = sum(L * 2 * (p + 1) * d * 2 for p in range(steps))
    total = steps * (L * w_layer + cross_kv) + max_new * logits + self_kv
    return {"per_token_step": L * w_layer + cross_kv + logits, "per_launch": total, "steps": steps,
            # B sessions per launch share the weight and logits streams; cross-KV and self-KV are per session
            "per_launch_batched": lambda B: steps * (L * w_layer + B * cross_kv) + max_new * logits + B * self_kv}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# =============================================================================================== reference arm
def build_hf_whisper(g):
    import torch
    from transformers import WhisperConfig, WhisperForConditionalGeneration
    cfg = WhisperConfig(vocab_size=g.vocab, num_mel_bins=g.n_mels, d_model=g.d_model, encoder_layers=g.enc_layers,
                        decoder_layers=g.dec_layers, encoder_attention_heads=g.heads, decoder_attention_heads=g.heads,
                        encoder_ffn_dim=g.ffn, decoder_ffn_dim=g.ffn, max_source_positions=g.max_source_positions,
                        max_target_positions=g.max_target_positions, pad_token_id=0, bos_token_id=0, eos_token_id=0,
                        decoder_start_token_id=PREFIX[0], suppress_tokens=None, begin_suppress_tokens=None)
    torch.manual_seed(0)
    return WhisperForConditionalGeneration(cfg).eval()


def run_reference_once(model, fe, audio):
    """What WhisperSTTHandler.process does on device='cpu', torch_dtype float32 (S/STT/whisper_stt_handler.py:83-87, 243)."""
    import torch
    feats = fe(audio, sampling_rate=16000, return_tensors="pt").input_features
    with torch.no_grad():
        out = model.generate(feats, decoder_input_ids=torch.tensor([PREFIX]), max_new_tokens=MAX_NEW,
                             min_new_tokens=MAX_NEW, num_beams=1, do_sample=False, suppress_tokens=SUPPRESS,
                             begin_suppress_tokens=BEGIN_SUPPRESS, return_timestamps=False)
    return out


def reference_arm(args, rank):
    """--impl reference: the transformers CPU path (the code the reference handler executes) on the host cores."""
    if rank != 0:
        return
    import torch
    from oracle import weights as W
    g = W.WHISPER_GEOMETRIES[MODEL]
    line = {"metric": METRIC, "unit": "sessions", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference", "config": {"workload": f"whisper-{MODEL} encode + greedy decode, single 10 s utterance",
                                            "max_new_tokens": MAX_NEW, "audio_s": AUDIO_S}}
    try:
        from transformers import WhisperFeatureExtractor
        import logging
        logging.getLogger("transformers").setLevel(logging.ERROR)
        model = build_hf_whisper(g)
        fe = WhisperFeatureExtractor(feature_size=g.n_mels)
        kind, what = "reference", f"transformers {__import__('transformers').__version__} WhisperForConditionalGeneration.generate fp32"
        fn = lambda a: run_reference_once(model, fe, a)
    except Exception as e:  # transformers missing: time the numpy oracle port instead
        from oracle import whisper_ref as R
        w = W.make_whisper_weights(g, 0)
        kind, what = "port", f"numpy oracle port (transformers unavailable: {type(e).__name__})"
        fn = lambda a: R.transcribe_ids(w, g, a, PREFIX, MAX_NEW, -1, SUPPRESS, BEGIN_SUPPRESS)
    audio = W.synthetic_audio(0, N_SAMPLES)
    for _ in range(max(0, args.warmup)):
        fn(audio)
    times = []
    for i in range(args.steps):
        a = W.synthetic_audio(i, N_SAMPLES)
        t = time.perf_counter()
        fn(a)
        times.append(time.perf_counter() - t)
    ms = 1e3 * sum(times) / len(times)
    val = AUDIO_S / (ms / 1e3)
    cores = torch.get_num_threads()
    line.update({"value": val, "ms_per_step": ms,
                 "cpu_baseline": {"value": val, "unit": "sessions", "cores": cores, "kind": kind,
                                  "sample": f"{args.steps} utterances x ({what}), {os.cpu_count()} host cpus"},
                 "e2e": {"value": val, "unit": "sessions", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "latency_ms_p50": 1e3 * statistics.median(times)})
    print(json.dumps(line), flush=True)


# =============================================================================================== B200 arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-baseline-utts", type=int, default=2, help="utterances timed on the host CPU (rank 0, N=1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default=MODEL)
    ap.add_argument("--profile-region", action="store_true",
                    help="cudaProfilerStart/Stop around the device-timed region (use with ncu --profile-from-start off)")
    ap.add_argument("--no-turn", action="store_true", help="skip the full-turn (STT -> LLM -> TTS post-proc) breakdown")
    ap.add_argument("--batch", type=int, default=16, help="sessions per launch for the secondary 'batched' figure (0 = skip)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from oracle import weights as W
    from speech_to_speech_b200 import engine as E, shard

    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    g = W.WHISPER_GEOMETRIES[args.model]
    eng = E.WhisperEngine(g.to_dict(), dtype="float16", max_batch=max(1, args.batch), device=local_rank)
    eng.init_random(seed=1234)
    opts = E.WhisperDecodeOptions(prefix=PREFIX, eos_id=-1, max_new_tokens=MAX_NEW, suppress=SUPPRESS,
                                  begin_suppress=BEGIN_SUPPRESS)
    dev = f"cuda:{local_rank}"
    n_in = args.warmup + args.steps
    # every step gets its own utterance; each rank a disjoint shard of the session stream (weak scaling, no collective)
    sessions = shard.local_sessions(rank, world, world * min(n_in, 8))  # global session ids owned by this rank
    auds = [W.synthetic_audio(sid, N_SAMPLES) for sid in sessions]
    pcm_dev = [torch.from_numpy(a)[None].to(dev).contiguous() for a in auds]
    pinned = [torch.from_numpy(a).pin_memory() for a in auds]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step_device(i):
        eng.logmel(pcm_dev[i % len(pcm_dev)], [N_SAMPLES])
        eng.encode(1)
        return eng.decode(1, opts)

    for i in range(args.warmup):
        step_device(i)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region 1: device-resident inputs, CUDA events per step, L2 flushed between steps ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    E.launch_count(local_rank, reset=True)
    barrier()
    if args.profile_region:
        torch.cuda.cudart().cudaProfilerStart()
    t_wall = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)          # L2 flush (not timed)
        s, mid, e = ev[i]
        s.record()
        eng.logmel(pcm_dev[i % len(pcm_dev)], [N_SAMPLES])
        eng.encode(1)
        mid.record()
        eng.decode(1, opts)
        e.record()
    barrier()
    if args.profile_region:
        torch.cuda.cudart().cudaProfilerStop()
    wall_s = time.perf_counter() - t_wall
    launches = E.launch_count(local_rank, reset=True)
    step_ms = [s.elapsed_time(e) for s, _, e in ev]
    dec_ms = [m.elapsed_time(e) for _, m, e in ev]
    enc_ms = [s.elapsed_time(m) for s, m, _ in ev]
    total_ms = sum(step_ms)

    # ---- timed region 2: end to end through the public host API (pinned host PCM in, ids on host out) ----
    for i in range(3):
        eng.transcribe([pinned[i % len(pinned)].numpy()], opts)
    barrier()
    e2e_t = []
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize()
        t = time.perf_counter()
        ids = eng.transcribe([pinned[i % len(pinned)].numpy()], opts)
        e2e_t.append(time.perf_counter() - t)
    barrier()
    e2e_total = sum(e2e_t)

    # ---- secondary figure: B concurrent sessions per launch (the weight stream is shared by the batch) ----
    batched = None
    if args.batch > 1:
        Bn = args.batch
        pcm_b = torch.stack([pcm_dev[i % len(pcm_dev)][0] for i in range(Bn)]).contiguous()
        host_b = [pinned[i % len(pinned)].numpy() for i in range(Bn)]
        for _ in range(2):
            eng.logmel(pcm_b, [N_SAMPLES] * Bn); eng.encode(Bn); eng.decode(Bn, opts)
        torch.cuda.synchronize()
        nb_steps = max(3, args.steps // 4)
        evb = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(nb_steps)]
        barrier()
        for i in range(nb_steps):
            flush.fill_(i & 0xFF)
            evb[i][0].record()
            eng.logmel(pcm_b, [N_SAMPLES] * Bn); eng.encode(Bn)
            evb[i][2].record()
            eng.decode(Bn, opts)
            evb[i][1].record()
        barrier()
        b_ms = [a.elapsed_time(b) for a, b, _ in evb]
        b_dec_ms = sum(m.elapsed_time(b) for _, b, m in evb) / nb_steps
        tb = []
        for i in range(nb_steps):
            t = time.perf_counter(); eng.transcribe(host_b, opts); tb.append(time.perf_counter() - t)
        barrier()
        b_tot, b_e2e = shard.max_over_ranks([sum(b_ms), sum(tb)], device=dev)
        batched = {"batch_per_gpu": Bn, "steps": nb_steps, "ms_per_step": b_tot / nb_steps,
                   "value": world * Bn * AUDIO_S / (b_tot / nb_steps / 1e3),
                   "e2e_value": world * Bn * AUDIO_S / (b_e2e / nb_steps), "latency_ms_p50": statistics.median(b_ms),
                   "decode_ms": b_dec_ms,
                   "note": "same kernels, Bn utterances per launch; every session still gets its full 128-token decode"}
    clocks = sampler.stop() if rank == 0 else None

    total_ms, e2e_total = shard.max_over_ranks([total_ms, e2e_total], device=dev)  # slowest rank defines the job

    if rank == 0:
        ms_per_step = total_ms / args.steps
        value = shard.whole_job_sessions(world, AUDIO_S, ms_per_step)
        e2e_value = shard.whole_job_sessions(world, AUDIO_S, 1e3 * e2e_total / args.steps)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        nb = decode_bytes_per_launch(g, len(PREFIX), MAX_NEW)
        dec_avg_ms = sum(dec_ms) / len(dec_ms)
        achieved = nb["per_launch"] / 1e9 / (dec_avg_ms / 1e3)
        cluster_on = os.environ.get("S2S_WHISPER_CLUSTER", "1") != "0"
        kernel_name = ("whisper_decode_cluster_kernel (persistent, 8-CTA cluster per head, 1 launch per utterance)" if cluster_on
                       else "whisper_decode_kernel (persistent, 1 launch per utterance)")
        traffic = None  # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the committed ncu --set full capture
        tpath = os.path.join(ROOT, "profiles", "ncu_decode_traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            traffic = tj.get("cluster_b1" if cluster_on else "grid_b1", {}).get("dram_bytes_per_launch")
        if batched is not None:
            bb = nb["per_launch_batched"](batched["batch_per_gpu"])
            ach_b = bb / 1e9 / (batched["decode_ms"] / 1e3)
            batched["roofline"] = {"kernel": "whisper_decode_kernel (persistent, %d sessions per launch)" % batched["batch_per_gpu"],
                                   "bound": "hbm", "achieved": ach_b, "peak": peak, "unit": "GB/s", "frac": ach_b / peak,
                                   "algorithmic_bytes_per_launch": bb,
                                   "traffic": (json.load(open(tpath)).get("grid_b%d" % batched["batch_per_gpu"], {}).get("dram_bytes_per_launch")
                                               if os.path.exists(tpath) else None)}
        line = {
            "metric": METRIC, "value": value, "unit": "sessions", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
            "config": {"workload": f"whisper-{args.model} log-mel + encoder + greedy decode, single 10 s utterance (BASELINE configs[1])",
                       "max_new_tokens": MAX_NEW, "audio_s": AUDIO_S, "batch_per_gpu": 1, "parallelism": f"session-shard dp{world}",
                       "l2": "flushed between timed steps (256 MiB write); per-token weight stream 335 MB > 126 MB L2",
                       "weights": "seeded random-init, exact geometry"},
            "latency_ms_p50": statistics.median(step_ms), "e2e_latency_ms_p50": 1e3 * statistics.median(e2e_t),
            "stage_ms": {"logmel_encoder": sum(enc_ms) / len(enc_ms), "decode_128_tokens": dec_avg_ms},
            "wall_s_timed_region": wall_s,
            "e2e": {"value": e2e_value, "unit": "sessions", "h2d_bytes_per_step": N_SAMPLES * 4,
                    "d2h_bytes_per_step": MAX_NEW * 4 + 4},
            "gpu_launches": int(launches),
            "roofline": {"kernel": kernel_name, "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": nb["per_launch"], "peak_source": peak_src,
                         "share_of_step": dec_avg_ms / ms_per_step},
            "clocks": clocks,
            "batched": batched,
        }
        if world == 1 and not args.no_turn:
            try:
                line["turn"] = full_turn(eng, opts, pinned[0].numpy(), local_rank)
            except Exception as e:  # never lose the headline line because of the extra section
                line["turn"] = {"error": f"{type(e).__name__}: {e}"}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(g, args.cpu_baseline_utts)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def full_turn(whisper_eng, opts, audio, device):
    """BASELINE configs[2] breakdown on one GPU, one session: Whisper-small STT -> Llama-3-8B (64-token prompt, 128-token
    reply, bf16, random-init) -> TTS post-processing of a 640 ms codec chunk.  The Qwen3-TTS model itself is not built
    (DESIGN.md section 7), so 'first_audio' below excludes the TTS model latency and says so."""
    import torch
    from oracle import weights as W
    from speech_to_speech_b200 import engine as E
    from speech_to_speech_b200.handlers.qwen3_tts_postproc import TTSPostProcessor
    lg = W.LLAMA_GEOMETRIES["llama-3-8b"]
    llm = E.LlamaEngine(lg.to_dict(), dtype="bfloat16", max_sessions=1, max_positions=1024, max_prefill=512, device=device)
    llm.init_random(7)
    post = TTSPostProcessor(device)
    prompt = np.random.default_rng(0).integers(0, lg.vocab, 64).tolist()
    chunk24k = (0.2 * np.sin(np.arange(15360) * 0.05)).astype(np.float32)  # 8 codec frames x 1920 samples
    rows = []
    for it in range(4):
        t0 = time.perf_counter()
        whisper_eng.transcribe([audio], opts)
        t1 = time.perf_counter()
        llm.reset(0)
        nxt, _ = llm.prefill(0, prompt)
        first = int(nxt[0])  # D2H: first reply token available on the host
        t2 = time.perf_counter()
        ids, lens = llm.decode([0], nxt, 19)  # first sentence ~20 tokens (what the TTS stage needs to start)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        ids, lens = llm.decode([0], ids[:, -1].contiguous(), 108)
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        post(chunk24k)
        t5 = time.perf_counter()
        rows.append([1e3 * (b - a) for a, b in ((t0, t1), (t1, t2), (t2, t3), (t3, t4), (t4, t5))])
    r = np.median(np.asarray(rows[1:]), axis=0)
    llm.close()
    return {"config": "whisper-small STT (128 tok) -> llama-3-8b bf16 (64-tok prompt, 128-tok reply) -> tts post-proc (640 ms chunk); 1 session",
            "stt_ms": r[0], "llm_prefill_first_token_ms": r[1], "llm_first_sentence_20tok_ms": r[2], "llm_remaining_108tok_ms": r[3],
            "tts_postproc_chunk_ms": r[4], "audio_in_to_first_sentence_ms": r[0] + r[1] + r[2],
            "llm_decode_ms_per_token": (r[2] + r[3]) / 127.0,
            "note": "Qwen3-TTS talker/codec not built (upstream absent, parity unpinned): first-audio latency would add its TTFA"}


def cpu_baseline(g, n_utts):
    """Bounded CPU sample beside the GPU number: the transformers fp32 path on the host cores (rank 0, N=1)."""
    import torch
    from oracle import weights as W
    try:
        from transformers import WhisperFeatureExtractor
        import logging
        logging.getLogger("transformers").setLevel(logging.ERROR)
        model = build_hf_whisper(g)
        fe = WhisperFeatureExtractor(feature_size=g.n_mels)
        run_reference_once(model, fe, W.synthetic_audio(0, N_SAMPLES))  # warm-up
        ts = []
        for i in range(n_utts):
            a = W.synthetic_audio(100 + i, N_SAMPLES)
            t = time.perf_counter()
            run_reference_once(model, fe, a)
            ts.append(time.perf_counter() - t)
        sec = sum(ts) / len(ts)
        return {"value": AUDIO_S / sec, "unit": "sessions", "cores": torch.get_num_threads(), "kind": "reference",
                "sample": f"{n_utts} utterances, transformers WhisperForConditionalGeneration.generate fp32 on "
                          f"{torch.get_num_threads()} threads ({os.cpu_count()} cpus), {sec:.2f} s/utterance"}
    except Exception as e:
        return {"value": None, "unit": "sessions", "cores": 0, "kind": "port", "sample": f"failed: {type(e).__name__}: {e}"}


if __name__ == "__main__":
    main()
