#!/bin/bash
# round-2 GPU pass D: tests, pointwise roofline after the strip-transpose preprocessing kernel
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/d_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_tests.log
timeout 300 python tools/pointwise_bench.py --json gpurun_out/d_pointwise.json > gpurun_out/d_pointwise.log 2>&1
tail -5 gpurun_out/d_tests.log
cat gpurun_out/d_pointwise.log
