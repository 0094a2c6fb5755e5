#!/bin/bash
# Developer aid: ncu launch list of TWO 8-frame chunks of the TTS stage (frame loop + codec decoder + post-processing, both
# lanes), and a --set full capture of the fp16-activation tensor-core contraction of the codec decoder.
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/prof; mkdir -p $O
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --profile-tts-chunks 2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file $O/launches_tts_r2.csv \
  $B > $O/launches_tts_bench.log 2>&1
echo "tts launch list rc=$? lines=$(wc -l < $O/launches_tts_r2.csv)"
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv1d_tc16_kernel -s 20 -c 8 -f -o $O/ncu_conv1d_tc16 \
  $B > $O/cap_conv16.log 2>&1; echo "conv16 rc=$?"
ls -la $O | tail -4
