#!/bin/bash
# A/B of an environment knob: VAR=NINT_L2_PERSIST VALUES="0 60 100" bash tools/exp_env.sh
for v in ${VALUES:-0}; do
  echo "##### ${VAR}=$v"
  env ${VAR}=$v python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('   ms/step', d['ms_per_step'], ' '.join(f\"{k} {v['avg_launch_us']}\" for k,v in d['kernels'].items()))
"
done
