#!/bin/bash
# time-fused conv launches: correctness first (bounded), then A/B on the same box
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "time_fused" > gpurun_out/fa_tests.log 2>&1
rc=$?
echo "pytest rc=$rc" >> gpurun_out/fa_tests.log
tail -15 gpurun_out/fa_tests.log
[ $rc -ne 0 ] && exit 0
for rep in 1 2 3; do
  for f in 0 1; do
    echo "fuse=$f $(NINT_FUSE_STEPS=$f timeout 300 python tools/step_time.py --steps 150 --bank 2>&1 | tail -1)" | tee -a gpurun_out/fa_ab.log
  done
done
for f in 0 1; do
  echo "shipped fuse=$f $(NINT_FUSE_STEPS=$f timeout 300 python tools/step_time.py --steps 30 --shipped 2>&1 | tail -1)" | tee -a gpurun_out/fa_ab.log
done
