#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/u_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/u_tests.log
tail -3 gpurun_out/u_tests.log
: > gpurun_out/u_ab.log
for rep in 1 2 3; do
  timeout 200 python tools/step_time.py --steps 150 --bank >> gpurun_out/u_ab.log 2>&1
done
cat gpurun_out/u_ab.log
