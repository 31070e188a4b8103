#!/bin/bash
cd "$(dirname "$0")/../.."
for env in "NINT_FUSE_STEPS=0" "NINT_FUSE_STEPS=2" "NINT_FUSE_STEPS=3" "NINT_FUSE_STEPS=0 NINT_CLUSTER=1"; do
  echo "== $env"; env $env timeout 200 python tests/tools/fuzz_one.py 1 4 5 12 24 "70,64,48" "5,3,3" bf16 11119 2>&1 | grep -v "^   " | tail -12
done
