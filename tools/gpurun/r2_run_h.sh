#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
bash tools/ncu_capture_r2.sh r2a > gpurun_out/h_ncu.log 2>&1
tail -20 gpurun_out/h_ncu.log
for c in "1 4 3 26 15 48 3 tf32 4022" "3 2 5 18 5 192 1 tf32 4067" "2 1 33 7 11 16,32,192 3,1,5 tf32 4094"; do
  echo "== fuzz_one $c" >> gpurun_out/h_fuzz_one.log
  python tests/tools/fuzz_one.py $c >> gpurun_out/h_fuzz_one.log 2>&1
done
cat gpurun_out/h_fuzz_one.log
