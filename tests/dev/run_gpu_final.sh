#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/final; mkdir -p $O
bash tests/dev/run_gpu_profile2.sh
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 2>&1 | tail -8 | cut -c1-300 | tee $O/pytest.log
echo "=== smoke under ncu"; timeout 300 ncu --metrics gpu__time_duration.sum python -c 'import __graft_entry__ as g; g.smoke(); print("__SMOKE_OK__")' > $O/ncu_smoke.log 2>&1; echo "rc=$? ok=$(grep -c __SMOKE_OK__ $O/ncu_smoke.log)"
echo "=== bench default"
timeout 900 python bench.py > $O/bench_r2_n1.json 2> $O/bench.err; echo "rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final/bench_r2_n1.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"]), {k: round(v) for k,v in d["stage_ms"].items()}, "roof", round(d["roofline"]["frac"],3))
print("e2e", json.dumps(d.get("e2e"))[:1100]); print("lat", d.get("latency_ms_p50"), d.get("latency_ms_p50_single_session"), d.get("latency_ms_at_70pct_load"))
PY
grep "bench " $O/bench.err | tail -5; tail -2 $O/bench.err | cut -c1-300
