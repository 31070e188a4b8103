#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NINT_DEBUG_FLAGS=2048 timeout 300 python tools/fused_debug.py 32 12 90 144 > gpurun_out/fd_debug.log 2>&1
echo "rc=$?" >> gpurun_out/fd_debug.log
grep "abandoned\|rc=\|Error" gpurun_out/fd_debug.log; grep "waits for" gpurun_out/fd_debug.log | head -40
grep "waits for" gpurun_out/fd_debug.log | awk '{print $5, $9}' | sort | uniq -c | sort -rn | head -20
NINT_DEBUG_FLAGS=2048 timeout 300 python tools/step_time.py --steps 5 --warmup 2 2>&1 | tail -3
