#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NINT_DEBUG_FLAGS=2048 timeout 300 python tools/fused_debug.py > gpurun_out/fc_debug.log 2>&1
echo "rc=$?" >> gpurun_out/fc_debug.log
grep -c "waits for" gpurun_out/fc_debug.log; grep "abandoned\|rc=" gpurun_out/fc_debug.log
if grep -q " [1-9][0-9]* abandoned" gpurun_out/fc_debug.log; then head -30 gpurun_out/fc_debug.log; exit 0; fi
bash tools/r2_run_fa.sh
