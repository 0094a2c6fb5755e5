#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2g; mkdir -p $O
echo "=== bench default full"
timeout 900 python bench.py --steps 1 --warmup 3 > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2g/bench_full.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"]), {k: round(v) for k,v in d["stage_ms"].items()}, "roof", round(d["roofline"]["frac"],3))
print("e2e", json.dumps(d.get("e2e"))[:1100]); print("lat", d.get("latency_ms_p50"), d.get("latency_ms_p50_single_session"))
PY
grep "bench " $O/bench_full.err | tail -4; tail -2 $O/bench_full.err | cut -c1-300
bash tests/dev/run_gpu_profile.sh
