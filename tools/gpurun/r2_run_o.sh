#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
: > gpurun_out/o_ab.log
for rep in 1 2; do
 for pdl in 0 1; do
  NINT_PDL=$pdl timeout 200 python tools/step_time.py --shipped --steps 30 --warmup 5 >> gpurun_out/o_ab.log 2>&1
  NINT_PDL=$pdl timeout 200 python tools/step_time.py --shipped --graph --steps 30 --warmup 5 >> gpurun_out/o_ab.log 2>&1
 done
done
cat gpurun_out/o_ab.log
