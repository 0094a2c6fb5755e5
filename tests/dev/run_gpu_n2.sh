#!/bin/bash
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/n2; mkdir -p $O
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --no-e2e > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?"
tail -c 1500 $O/bench_n2.json; grep "bench \|Error\|error" $O/bench_n2.err | tail -8 | cut -c1-300
