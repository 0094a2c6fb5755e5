#!/bin/bash
# Developer aid: GPU test suite + the ncu-wrapped smoke the driver runs + the cluster-attribute probe.
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/round; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 --timeout=900 2>&1 | tail -40 | tee $O/pytest.log
echo "=== cluster probe (plain / under ncu)"
tests/dev/cluster_probe 2>&1 | tee $O/probe_plain.log
ncu --metrics gpu__time_duration.sum tests/dev/cluster_probe 2>&1 | grep -v "^==PROF==" | grep "coop=" | tee $O/probe_ncu.log
echo "=== ncu-wrapped smoke (cluster on / off)"
for cl in 1 0; do
  S2S_WHISPER_CLUSTER=$cl ncu --metrics gpu__time_duration.sum python -c 'import __graft_entry__ as g; g.smoke(); print("__SMOKE_OK__")' > $O/ncu_smoke_cl$cl.log 2>&1
  echo "cluster=$cl rc=$? $(grep -c __SMOKE_OK__ $O/ncu_smoke_cl$cl.log) ok; $(grep -h 'smoke:' $O/ncu_smoke_cl$cl.log | tail -1 | cut -c1-220)"
done
