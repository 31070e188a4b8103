#!/bin/bash
# 8-GPU regression of the final build: the driver's own command
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/g8_bench.json 2> gpurun_out/g8_bench.err
echo "bench rc=$?"
grep '^{' gpurun_out/g8_bench.json | tail -1 | cut -c1-400
