#!/bin/bash
# Developer aid: the round's ncu evidence.  1) launch list of the bench's device-timed region; 2) --set full captures of the
# dominant kernels (one GPU; durations under ncu are never bench numbers).
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/prof; mkdir -p $O
S=${SESSIONS:-16}
timeout 1700 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 60000 --csv --log-file $O/launches_r2.csv \
  python bench.py --steps 1 --warmup 3 --sessions $S --no-e2e --no-cpu-baseline --profile-region > $O/launches_bench.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $O/launches_r2.csv)"
# full captures: llama decode (the 2nd launch inside the profiled region), the largest tensor-core contraction, whisper decode, the talker step
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:llama_decode_kernel -s 1 -c 1 -f -o $O/ncu_llama_decode \
  python bench.py --steps 1 --warmup 3 --sessions $S --no-e2e --no-cpu-baseline --profile-region > $O/cap_llama.log 2>&1; echo "llama rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:whisper_decode_kernel -c 1 -f -o $O/ncu_whisper_decode \
  python bench.py --steps 1 --warmup 3 --sessions $S --no-e2e --no-cpu-baseline --profile-region > $O/cap_whisper.log 2>&1; echo "whisper rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv1d_tc_kernel -s 400 -c 6 -f -o $O/ncu_conv1d_tc \
  python bench.py --steps 1 --warmup 3 --sessions $S --no-e2e --no-cpu-baseline --profile-region > $O/cap_conv.log 2>&1; echo "conv rc=$?"
ls -la $O
