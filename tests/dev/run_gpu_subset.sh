#!/bin/bash
# Developer aid: a subset of the GPU tests + the ncu-wrapped smoke.
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/subset; mkdir -p $O
timeout 1200 python -m pytest "$@" -m gpu -q --timeout=900 2>&1 | tail -60 | tee $O/pytest.log
for cl in 1; do
  S2S_WHISPER_CLUSTER=$cl ncu --metrics gpu__time_duration.sum python -c 'import __graft_entry__ as g; g.smoke(); print("__SMOKE_OK__")' > $O/ncu_smoke_cl$cl.log 2>&1
  echo "ncu smoke cluster=$cl rc=$? $(grep -c __SMOKE_OK__ $O/ncu_smoke_cl$cl.log) ok; $(grep -h 'smoke:' $O/ncu_smoke_cl$cl.log | tail -1 | cut -c1-220)"
done
