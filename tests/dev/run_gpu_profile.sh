#!/bin/bash
# Developer aid: the round's ncu evidence.  1) launch list of the bench's device-timed region; 2) --set full captures of the
# dominant kernels (one GPU; durations under ncu are never bench numbers).
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/prof; mkdir -p $O
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --profile-region"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 90000 --csv --log-file $O/launches_r2.csv \
  $B > $O/launches_bench.log 2>&1
echo "launch list rc=$? lines=$(wc -l < $O/launches_r2.csv)"
# full captures: Llama-3-8B decode (first llama_decode_kernel launches of the region), then a talker + a predictor launch of the TTS
# frame loop (the same kernel, later in the region), the Whisper decode kernel, and the largest tensor-core contractions of the codec
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:llama_decode_kernel -s 1 -c 1 -f -o $O/ncu_llama_decode \
  $B > $O/cap_llama.log 2>&1; echo "llama rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:llama_decode_kernel -s 8 -c 2 -f -o $O/ncu_tts_decode \
  $B > $O/cap_tts.log 2>&1; echo "tts rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:whisper_decode_kernel -c 1 -f -o $O/ncu_whisper_decode \
  $B > $O/cap_whisper.log 2>&1; echo "whisper rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv1d_tc_kernel -s 400 -c 8 -f -o $O/ncu_conv1d_tc \
  $B > $O/cap_conv.log 2>&1; echo "conv rc=$?"
ls -la $O
