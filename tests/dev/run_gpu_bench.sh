#!/bin/bash
# Developer aid: selected GPU tests, then the bench (both arms).
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/bench; mkdir -p $O
if [ -n "$1" ]; then timeout 1200 python -m pytest $1 -m gpu -q --timeout=900 2>&1 | tail -40 | tee $O/pytest.log; fi
echo "=== bench b200"
timeout 1500 python bench.py --gpus 1 --steps ${STEPS:-2} --warmup 3 --sessions ${SESSIONS:-16} > $O/bench.json 2> $O/bench.err; echo "rc=$?"; tail -c 6000 $O/bench.json; tail -20 $O/bench.err
echo "=== bench reference"
timeout 900 python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?"; tail -c 3000 $O/bench_ref.json; tail -5 $O/bench_ref.err
