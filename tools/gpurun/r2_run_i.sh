#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for f in 8 9 10; do
  echo "=== bwd NINT_DEBUG_FLAGS=$f" >> gpurun_out/i_trace.log
  NINT_DEBUG_FLAGS=$f timeout 120 python tools/trace_report.py bwd >> gpurun_out/i_trace.log 2>&1
done
for f in 8 9; do
  echo "=== fwd NINT_DEBUG_FLAGS=$f" >> gpurun_out/i_trace.log
  NINT_DEBUG_FLAGS=$f timeout 120 python tools/trace_report.py fwd >> gpurun_out/i_trace.log 2>&1
done
grep -E "^===|steady period|== role" gpurun_out/i_trace.log
timeout 600 python -m pytest tests -m gpu -q -x -k "head or golden or return_sequence or val_loop or sensitivity" > gpurun_out/i_tests.log 2>&1
tail -3 gpurun_out/i_tests.log
