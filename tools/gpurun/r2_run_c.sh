#!/bin/bash
# round-2 GPU pass C: tests again (pointwise kernels changed), pointwise roofline, PDL / bank / graph A-B without profiling events
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_tests.log
timeout 300 python tools/pointwise_bench.py --json gpurun_out/c_pointwise.json > gpurun_out/c_pointwise.log 2>&1
: > gpurun_out/c_ab.log
for rep in 1 2; do
  for pdl in 0 1; do
    NINT_PDL=$pdl timeout 200 python tools/step_time.py --steps 150 >> gpurun_out/c_ab.log 2>&1
    NINT_PDL=$pdl timeout 200 python tools/step_time.py --steps 150 --bank >> gpurun_out/c_ab.log 2>&1
  done
done
NINT_PDL=0 timeout 200 python tools/step_time.py --steps 150 --bank --graph >> gpurun_out/c_ab.log 2>&1
NINT_PDL=1 timeout 200 python tools/step_time.py --steps 150 --bank --graph >> gpurun_out/c_ab.log 2>&1
NINT_PDL=0 timeout 200 python tools/step_time.py --steps 300 --batch 2 --seq-len 4 >> gpurun_out/c_ab.log 2>&1
NINT_PDL=0 timeout 200 python tools/step_time.py --steps 300 --batch 2 --seq-len 4 --graph >> gpurun_out/c_ab.log 2>&1
NINT_PDL=1 timeout 200 python tools/step_time.py --steps 300 --batch 2 --seq-len 4 --graph >> gpurun_out/c_ab.log 2>&1
tail -3 gpurun_out/c_tests.log
cat gpurun_out/c_pointwise.log gpurun_out/c_ab.log
