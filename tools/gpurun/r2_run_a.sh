#!/bin/bash
# round-2 GPU pass A: full GPU test suite, default bench line, sub-batch-major schedule experiment
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
lscpu | head -20 >> gpurun_out/a_gpu.txt 2>&1
nvidia-smi topo -m >> gpurun_out/a_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q -s --durations=15 > gpurun_out/a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_tests.log
timeout 600 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?" >> gpurun_out/a_bench.err
for sb in 16 8 4; do
  NINT_SUB_BATCH=$sb timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/a_bench_sb$sb.json 2> gpurun_out/a_bench_sb$sb.err
done
tail -5 gpurun_out/a_tests.log
cat gpurun_out/a_bench.json | cut -c1-600
