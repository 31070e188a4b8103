#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/w_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/w_tests.log
tail -3 gpurun_out/w_tests.log
: > gpurun_out/w_ab.log
for rep in 1 2 3; do
  for f in 1024 0; do
    echo "flags=$f" >> gpurun_out/w_ab.log
    NINT_DEBUG_FLAGS=$f timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 60 --warmup 5 2>/dev/null | python -c "
import json,sys
e=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print(e['value'], e['ms_per_step'], {k:(v['ms_per_step'],v['avg_launch_us']) for k,v in e['kernels'].items()}, e['clocks']['power_w'])" >> gpurun_out/w_ab.log
  done
done
cat gpurun_out/w_ab.log
