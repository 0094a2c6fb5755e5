#!/bin/bash
# Developer aid: reproduce the round-1 driver failure (ncu-wrapped smoke returned wrong ids) and run the sanitizers.
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/diag; mkdir -p $O

SMOKE='import __graft_entry__ as g; g.smoke(); print("__SMOKE_OK__")'
run() { name=$1; shift; echo "=== $name"; ( "$@" ) > $O/$name.log 2>&1; echo "rc=$? $(grep -c __SMOKE_OK__ $O/$name.log) ok; $(grep -h 'smoke:' $O/$name.log | tail -1 | cut -c1-200)"; }
run plain python -c "$SMOKE"
for cl in 1 0; do for rx in 0 1; do
  run ncu_default_cl${cl}_rx${rx} env S2S_WHISPER_CLUSTER=$cl S2S_SYNC_RELAXED=$rx ncu --metrics gpu__time_duration.sum python -c "$SMOKE"
done; done
run ncu_nocontrol_cl1_rx1 env S2S_SYNC_RELAXED=1 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none python -c "$SMOKE"
run ncu_noclock_cl1_rx1 env S2S_SYNC_RELAXED=1 ncu --metrics gpu__time_duration.sum --clock-control none python -c "$SMOKE"
run ncu_nocache_cl1_rx1 env S2S_SYNC_RELAXED=1 ncu --metrics gpu__time_duration.sum --cache-control none python -c "$SMOKE"
for tool in memcheck initcheck racecheck synccheck; do
  for cl in 1 0; do
    run san_${tool}_cl${cl} env S2S_WHISPER_CLUSTER=$cl timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -c "$SMOKE"
  done
done
echo "=== sync A/B timing"; timeout 600 python tests/dev/sync_ab.py 2>&1 | tee $O/sync_ab.log
