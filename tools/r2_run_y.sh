#!/bin/bash
# memcheck of the round-2 kernels on small cases (one compute-sanitizer tool per call)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "bf16_input or frame_bank or fused_preprocessing or step_windows or cell_is or input_gradient or arbitrary_hidden or second_backward or deterministic" \
  > gpurun_out/y_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/y_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/y_memcheck.log
tail -8 gpurun_out/y_memcheck.log
