#!/bin/bash
# round-2 GPU pass B: full GPU test suite (no -x), PDL A/B, pointwise roofline
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s --durations=20 > gpurun_out/b_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_tests.log
for pdl in 0 1 0 1; do
  NINT_PDL=$pdl timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 40 --warmup 5 >> gpurun_out/b_bench_pdl$pdl.json 2>> gpurun_out/b_bench_pdl.err
done
NINT_PDL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/b_tests_pdl.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_tests_pdl.log
timeout 300 python tools/pointwise_bench.py --json gpurun_out/b_pointwise.json > gpurun_out/b_pointwise.log 2>&1
tail -15 gpurun_out/b_tests.log
tail -3 gpurun_out/b_tests_pdl.log
cat gpurun_out/b_pointwise.log
