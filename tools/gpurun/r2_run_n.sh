#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for rep in 1 2; do
timeout 300 python bench.py --precision tf32 --no-extras --no-cpu-baseline --steps 20 --warmup 5 2> gpurun_out/n_bench_tf32.err | grep "^{" > gpurun_out/n_bench_tf32.json
python -c "
import json
e=json.load(open('gpurun_out/n_bench_tf32.json'))
print('tf32', e['value'], e['ms_per_step'], {k:(v['ms_per_step'],v['avg_launch_us']) for k,v in e['kernels'].items()})"
done
timeout 900 python -m pytest tests -m gpu -q -x -k "tf32 or determin or cell or hidden" > gpurun_out/n_tests.log 2>&1; tail -2 gpurun_out/n_tests.log
timeout 600 python tests/tools/fuzz_parity.py 160 4 > gpurun_out/n_fuzz.log 2>&1
tail -1 gpurun_out/n_fuzz.log; grep "marg\|FAIL\|ERROR" gpurun_out/n_fuzz.log | head
