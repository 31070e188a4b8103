#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for f in 0 2 3 2 0; do
  echo "fuse=$f $(NINT_FUSE_STEPS=$f timeout 300 python tools/fused_stress.py 400 --det 2>&1 | grep 'steps ok\|code=[1-9]\|Error')"
done
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "time_fused" 2>&1 | tail -2
