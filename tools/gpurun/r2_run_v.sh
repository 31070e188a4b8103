#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
: > gpurun_out/v_ab.log
for rep in 1 2 3; do
  echo "prev build" >> gpurun_out/v_ab.log
  NINT_LIB=$PWD/tools/ab/libnint_prev.so timeout 200 python tools/step_time.py --steps 150 --bank >> gpurun_out/v_ab.log 2>&1
  echo "current build" >> gpurun_out/v_ab.log
  timeout 200 python tools/step_time.py --steps 150 --bank >> gpurun_out/v_ab.log 2>&1
done
cat gpurun_out/v_ab.log
timeout 900 python -m pytest tests -m gpu -q -x -k "golden or rollout or plans or edge or cfg2 or linearity" > gpurun_out/v_tests.log 2>&1; tail -2 gpurun_out/v_tests.log
