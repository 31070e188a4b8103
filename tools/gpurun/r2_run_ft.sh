#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
for rep in 1 2; do
for f in 0 2 3 auto; do
  if [ $f = auto ]; then unset NINT_FUSE_STEPS; else export NINT_FUSE_STEPS=$f; fi
  echo "shipped fuse=$f $(timeout 300 python tools/step_time.py --steps 30 --shipped 2>&1 | tail -1)"
done
done
for f in 0 auto; do
  if [ $f = auto ]; then unset NINT_FUSE_STEPS; else export NINT_FUSE_STEPS=$f; fi
  echo "B8 fuse=$f $(timeout 300 python tools/step_time.py --steps 100 --batch 8 2>&1 | tail -1)"
  echo "B32 fuse=$f $(timeout 300 python tools/step_time.py --steps 100 2>&1 | tail -1)"
done
