#!/bin/bash
# Developer aid: selected GPU tests, then the bench in stages (device-timed first, then the handler-level e2e, then the CPU arm).
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/bench; mkdir -p $O
if [ -n "$1" ]; then timeout 900 python -m pytest $1 -m gpu -q -x --timeout=600 2>&1 | tail -30 | tee $O/pytest.log; fi
echo "=== bench b200 (device-timed only)"
timeout 420 python bench.py --gpus 1 --steps ${STEPS:-1} --warmup 3 --sessions ${SESSIONS:-16} --no-e2e --no-cpu-baseline > $O/bench_dev.json 2> $O/bench_dev.err; echo "rc=$?"; tail -c 5000 $O/bench_dev.json; tail -12 $O/bench_dev.err
if [ "${E2E:-1}" = "1" ]; then
echo "=== bench b200 (full line)"
timeout 900 python bench.py --gpus 1 --steps ${STEPS:-1} --warmup 3 --sessions ${SESSIONS:-16} > $O/bench.json 2> $O/bench.err; echo "rc=$?"; tail -c 7000 $O/bench.json; tail -25 $O/bench.err
fi
if [ "${REF:-0}" = "1" ]; then
echo "=== bench reference"
timeout 400 python bench.py --impl reference --gpus 1 --steps 1 --warmup 0 > $O/bench_ref.json 2> $O/bench_ref.err; echo "rc=$?"; tail -c 3000 $O/bench_ref.json; tail -5 $O/bench_ref.err
fi
