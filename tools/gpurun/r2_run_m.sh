#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for c in "3 2 5 22 33 70,16,192 3,3,3 tf32 4024" "1 1 3 18 27 48,24,70 3,3,1 tf32 4134"; do
  echo "== fuzz_one $c" >> gpurun_out/m_fuzz_one.log
  python tests/tools/fuzz_one.py $c >> gpurun_out/m_fuzz_one.log 2>&1
done
cat gpurun_out/m_fuzz_one.log
timeout 300 python bench.py --precision tf32 --no-extras --no-cpu-baseline --steps 20 --warmup 5 2> gpurun_out/m_bench_tf32.err | grep "^{" > gpurun_out/m_bench_tf32.json
python -c "
import json
e=json.load(open('gpurun_out/m_bench_tf32.json'))
print('tf32', e['value'], e['ms_per_step'], {k:(v['ms_per_step'],v['avg_launch_us']) for k,v in e['kernels'].items()})"
timeout 600 python -m pytest tests -m gpu -q -x -k "tf32" > gpurun_out/m_tests.log 2>&1; tail -2 gpurun_out/m_tests.log
