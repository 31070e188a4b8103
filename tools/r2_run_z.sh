#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tests/tools/fuzz_parity.py 300 7 > gpurun_out/z_fuzz.log 2>&1
echo "fuzz rc=$?" >> gpurun_out/z_fuzz.log
tail -2 gpurun_out/z_fuzz.log; grep "marg\|FAIL\|ERROR\|refused" gpurun_out/z_fuzz.log | head -30
