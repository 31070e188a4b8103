#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for f in 0 2; do
  NINT_DEBUG_FLAGS=8 NINT_FUSE_STEPS=$f timeout 300 python tools/trace_report.py bwd > gpurun_out/fm_trace_bwd_fuse$f.log 2>&1
  echo "== fuse=$f"; grep "== role\|steady" gpurun_out/fm_trace_bwd_fuse$f.log
done
