#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "fuse or preprocessing or shipped" > gpurun_out/f_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_tests.log
timeout 300 python tools/pointwise_bench.py --json gpurun_out/f_pointwise.json > gpurun_out/f_pointwise.log 2>&1
tail -4 gpurun_out/f_tests.log
head -3 gpurun_out/f_pointwise.log
