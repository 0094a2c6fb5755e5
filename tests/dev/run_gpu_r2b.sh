#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2b; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_codec.py tests/test_gpu_qwen3tts.py -m gpu -q --timeout=600 2>&1 | tail -15 | tee $O/pytest.log
echo "=== lanes"; timeout 400 python tests/dev/dev_lanes.py 2>&1 | tee $O/lanes.txt | tail -12
echo "=== trace B=16 (codec chunk time after trimming)"; timeout 300 python tests/dev/dev_trace_tts.py 16 2>&1 | tail -4 | tee $O/trace16.txt
