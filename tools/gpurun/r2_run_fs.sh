#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for rep in 1 2; do
for f in 0 2 3; do
  echo "shipped fuse=$f $(NINT_FUSE_STEPS=$f timeout 300 python tools/step_time.py --steps 30 --shipped 2>&1 | tail -1)"
done
done
for f in 0 2 3; do
  echo "B8 fuse=$f $(NINT_FUSE_STEPS=$f timeout 300 python tools/step_time.py --steps 100 --batch 8 2>&1 | tail -1)"
  echo "B2T4 fuse=$f $(NINT_FUSE_STEPS=$f timeout 300 python tools/step_time.py --steps 300 --batch 2 --seq-len 4 2>&1 | tail -1)"
  echo "B16 fuse=$f $(NINT_FUSE_STEPS=$f timeout 300 python tools/step_time.py --steps 100 --batch 16 2>&1 | tail -1)"
done
