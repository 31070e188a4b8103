#!/bin/bash
# final single-GPU regression: build check, full GPU suite, smoke, default bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/p_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p_tests.log
tail -4 gpurun_out/p_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/p_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/p_smoke.log; tail -3 gpurun_out/p_smoke.log
timeout 900 python bench.py > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; echo "bench rc=$?"
tail -1 gpurun_out/p_bench.json | cut -c1-300
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/p_bench_ref.json 2>> gpurun_out/p_bench.err
cut -c1-200 gpurun_out/p_bench_ref.json
