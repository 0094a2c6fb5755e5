#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2a; mkdir -p $O
echo "=== TTS phase trace B=16"; timeout 300 python tests/dev/dev_trace_tts.py 16 2>&1 | tee $O/trace_tts_b16.txt | tail -40
echo "=== TTS phase trace B=1"; timeout 300 python tests/dev/dev_trace_tts.py 1 2>&1 | tee $O/trace_tts_b1.txt | tail -30
SMOKE='import __graft_entry__ as g; g.smoke(); print("__SMOKE_OK__")'
for cl in 1 0; do
  S2S_WHISPER_CLUSTER=$cl timeout 600 compute-sanitizer --tool racecheck --print-limit 30 python -c "$SMOKE" > $O/racecheck_cl$cl.log 2>&1
  echo "racecheck cluster=$cl: $(grep -c __SMOKE_OK__ $O/racecheck_cl$cl.log) ok; $(grep 'RACECHECK SUMMARY' $O/racecheck_cl$cl.log)"
done
echo "=== bench full line"
timeout 700 python bench.py --gpus 1 --steps 1 --warmup 3 --sessions 16 > $O/bench.json 2> $O/bench.err; echo "rc=$?"; tail -c 6000 $O/bench.json; grep "bench " $O/bench.err | tail -12
