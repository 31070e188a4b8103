#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
(cd ab_old && NINT_DEBUG_FLAGS=8 timeout 300 python tools/trace_report.py fwd > ../gpurun_out/fq_trace_old.log 2>&1)
NINT_FUSE_STEPS=0 NINT_DEBUG_FLAGS=8 timeout 300 python tools/trace_report.py fwd > gpurun_out/fq_trace_new.log 2>&1
for f in old new; do echo "== $f"; grep "== role\|steady" gpurun_out/fq_trace_$f.log; done
