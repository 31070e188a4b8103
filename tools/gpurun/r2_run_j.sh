#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
NINT_DEBUG_FLAGS=256 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/j_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j_tests.log
tail -3 gpurun_out/j_tests.log
: > gpurun_out/j_ab.log
for rep in 1 2; do
  for cfg in "0 0" "256 0" "256 2"; do
    set -- $cfg
    echo "flags=$1 plan_ns=$2" >> gpurun_out/j_ab.log
    NINT_DEBUG_FLAGS=$1 NINT_PLAN_NS=$2 timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 40 --warmup 5 2>> gpurun_out/j_ab.err | python -c "
import json,sys
e=json.loads(sys.stdin.readline())
print(e['value'], e['ms_per_step'], {k:(v['ms_per_step'],v['avg_launch_us']) for k,v in e['kernels'].items()})" >> gpurun_out/j_ab.log
  done
done
cat gpurun_out/j_ab.log
for f in 264; do
  echo "=== bwd NINT_DEBUG_FLAGS=$f" >> gpurun_out/j_trace.log
  NINT_DEBUG_FLAGS=$f timeout 120 python tools/trace_report.py bwd >> gpurun_out/j_trace.log 2>&1
done
grep -E "^===|steady period|== role" gpurun_out/j_trace.log
