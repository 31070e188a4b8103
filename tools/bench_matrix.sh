#!/bin/bash
# A/B matrix of kernel variants; prints the per-kernel breakdown of bench.py for each
for cfg in ${CFGS:-"pertap 1" "halo 1" "halo 2"}; do
  set -- $cfg
  echo "##### variant=$1 cluster=$2"
  NINT_CONV_VARIANT=$1 NINT_CLUSTER=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
for k,v in d['kernels'].items(): print('  ', k, v)
print('  ', d['gate_conv_fwd_bwd'])
"
done
