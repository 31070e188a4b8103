#!/bin/bash
# the automatic time-fusing rule re-checked on the last build (one-launch-per-step kernels got faster)
cd "$(dirname "$0")/../.."
for rep in 1 2; do
for f in 0 auto 3; do
  if [ $f = auto ]; then unset NINT_FUSE_STEPS; else export NINT_FUSE_STEPS=$f; fi
  echo "shipped fuse=$f $(timeout 300 python tools/step_time.py --steps 30 --shipped 2>&1 | tail -1)"
done
done
for b in 8 16 32; do
for f in 0 2; do
  export NINT_FUSE_STEPS=$f
  echo "B$b fuse=$f $(timeout 300 python tools/step_time.py --steps 100 --batch $b 2>&1 | tail -1)"
done
done
