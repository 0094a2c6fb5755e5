#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2f; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_codec.py tests/test_gpu_qwen3tts.py -m gpu -q -s --timeout=600 2>&1 | grep -E "vs fp16|passed|failed|Error|error" | cut -c1-400 | tee $O/pytest_codec.log
echo "=== trace"; timeout 300 python tests/dev/dev_trace_tts.py 16 2>&1 | tail -4
echo "=== bench default full"
timeout 900 python bench.py --steps 1 --warmup 3 > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2f/bench_full.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"]), {k: round(v) for k,v in d["stage_ms"].items()}, "roof", round(d["roofline"]["frac"],3))
print("e2e", json.dumps(d.get("e2e"))[:1100]); print("lat", d.get("latency_ms_p50"), d.get("latency_ms_p50_single_session"))
PY
grep "bench " $O/bench_full.err | tail -6; tail -2 $O/bench_full.err | cut -c1-300
