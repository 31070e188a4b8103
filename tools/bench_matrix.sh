#!/bin/bash
# A/B of CTA pairs (tcgen05 cta_group::2) against single CTAs; prints the per-kernel breakdown of bench.py for each
for cl in ${CLUSTERS:-2 1}; do
  echo "##### NINT_CLUSTER=$cl"
  NINT_CLUSTER=$cl python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])
for k,v in d['kernels'].items(): print('  ', k, v)
print('  ', d['gate_conv_fwd_bwd'])
"
done
