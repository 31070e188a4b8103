#!/bin/bash
cd "$(dirname "$0")/../.."
for b in 8 16 32; do
for f in 0 1; do
  export NINT_FUSE_STEPS=$f
  echo "B$b fuse=$f $(timeout 300 python tools/step_time.py --steps 100 --batch $b 2>&1 | tail -1)"
done
done
for f in 0 2; do
  export NINT_FUSE_STEPS=$f
  echo "B24 fuse=$f $(timeout 300 python tools/step_time.py --steps 100 --batch 24 2>&1 | tail -1)"
done
for f in 0 1 0 1; do
  export NINT_FUSE_STEPS=$f
  echo "B8 fuse=$f $(timeout 300 python tools/step_time.py --steps 100 --batch 8 2>&1 | tail -1)"
done
