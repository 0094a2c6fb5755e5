#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/last; mkdir -p $O
timeout 330 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $O/bench.json 2> $O/bench.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/last/bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"]))
print("e2e", json.dumps(d.get("e2e"))[:1100]); print("lat", d.get("latency_ms_p50"), d.get("latency_ms_p50_single_session"), d.get("latency_ms_at_70pct_load"))
PY
grep "bench " $O/bench.err | tail -5; tail -2 $O/bench.err | cut -c1-300
