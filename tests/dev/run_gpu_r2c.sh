#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2c; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_codec.py -m gpu -q --timeout=600 -k real_geometry 2>&1 | tail -5 | tee $O/pytest_codec.log
echo "=== bench lanes=2 S=32 (device only)"
timeout 900 python bench.py --steps 1 --warmup 3 --sessions 32 --lanes 2 --no-e2e --no-cpu-baseline > $O/bench_l2.json 2> $O/bench_l2.err; echo "rc=$?"; tail -c 3500 $O/bench_l2.json; grep "bench " $O/bench_l2.err | tail -8; tail -3 $O/bench_l2.err
echo "=== bench lanes=1 S=16 (device only)"
timeout 600 python bench.py --steps 1 --warmup 3 --sessions 16 --lanes 1 --no-e2e --no-cpu-baseline > $O/bench_l1.json 2> $O/bench_l1.err; echo "rc=$?"; python - <<'PY'
import json
for n in ("l2","l1"):
    try:
        d=json.loads(open(f"gpurun_out/r2c/bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "value", round(d["value"],1), "ms/step", round(d["ms_per_step"]), {k: round(v) for k,v in d["stage_ms"].items()}, "roof", round(d["roofline"]["frac"],3))
    except Exception as e: print(n, "ERR", e)
PY
echo "=== bench lanes=2 full (e2e + cpu)"
timeout 900 python bench.py --steps 1 --warmup 3 > $O/bench_full.json 2> $O/bench_full.err; echo "rc=$?"; tail -c 2500 $O/bench_full.json; grep "bench " $O/bench_full.err | tail -12; tail -3 $O/bench_full.err
