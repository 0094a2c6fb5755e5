#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2d; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_llama.py tests/test_gpu_qwen3tts.py tests/test_gpu_codec.py tests/test_gpu_handlers.py -m gpu -q -s --timeout=600 2>&1 | grep -v "^$" | tail -40 | cut -c1-400 | tee $O/pytest.log
echo "=== TTS trace B=16"; timeout 300 python tests/dev/dev_trace_tts.py 16 2>&1 | tee $O/trace_tts_b16.txt | tail -24
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"]), {k: round(v) for k,v in d["stage_ms"].items()}, "roof", round(d["roofline"]["frac"],3), "e2e", (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("errors"), "lat", d.get("latency_ms_p50"), d.get("latency_ms_p50_single_session"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
}
for L in 2 3 4; do
  timeout 600 python bench.py --steps 1 --warmup 3 --sessions $((16*L)) --lanes $L --no-e2e --no-cpu-baseline > $O/bench_l$L.json 2> $O/bench_l$L.err; echo "lanes $L rc=$?"; summ $O/bench_l$L.json; tail -2 $O/bench_l$L.err | cut -c1-300
done
echo "=== bench default full"
timeout 900 python bench.py --steps 1 --warmup 3 > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?"; summ $O/bench_full.json; grep "bench " $O/bench_full.err | tail -12; tail -3 $O/bench_full.err | cut -c1-400
