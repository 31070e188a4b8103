#!/bin/bash
# 2-GPU regression of the final build: the two 2-GPU tests, then the default bench under torchrun
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "fused_nvlink or non_current_device" > gpurun_out/g2_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/g2_tests.log
tail -3 gpurun_out/g2_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/g2_bench.json 2> gpurun_out/g2_bench.err
echo "bench rc=$?"
grep '^{' gpurun_out/g2_bench.json | tail -1 | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | grep '^{' | cut -c1-300
