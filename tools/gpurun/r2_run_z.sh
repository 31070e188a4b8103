#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python tests/tools/fuzz_parity.py 300 11 > gpurun_out/z_fuzz_auto.log 2>&1
echo "fuzz auto rc=$?"; tail -1 gpurun_out/z_fuzz_auto.log; grep "FAIL\|ERROR\|BAD\|bad " gpurun_out/z_fuzz_auto.log | head
NINT_FUSE_STEPS=3 timeout 900 python tests/tools/fuzz_parity.py 300 12 > gpurun_out/z_fuzz_fuse3.log 2>&1
echo "fuzz fuse3 rc=$?"; tail -1 gpurun_out/z_fuzz_fuse3.log; grep "FAIL\|ERROR\|BAD\|bad " gpurun_out/z_fuzz_fuse3.log | head
