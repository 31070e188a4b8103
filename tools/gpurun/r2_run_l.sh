#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/l_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/l_tests.log
tail -15 gpurun_out/l_tests.log
timeout 600 python tests/tools/fuzz_parity.py 160 4 > gpurun_out/l_fuzz.log 2>&1
tail -3 gpurun_out/l_fuzz.log; grep -c "marg" gpurun_out/l_fuzz.log; grep "marg\|FAIL\|ERROR" gpurun_out/l_fuzz.log | head -20
for prec in tf32; do
  timeout 300 python bench.py --precision $prec --no-extras --no-cpu-baseline --steps 20 --warmup 5 2> gpurun_out/l_bench_$prec.err | grep "^{" > gpurun_out/l_bench_$prec.json
  python -c "
import json
e=json.load(open('gpurun_out/l_bench_$prec.json'))
print('$prec', e['value'], e['ms_per_step'], {k:(v['ms_per_step'],v['avg_launch_us']) for k,v in e['kernels'].items()})"
done
python __graft_entry__.py smoke 2>&1 | tail -3
