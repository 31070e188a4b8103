#!/bin/bash
# the kernel knobs exist only in the experiment build: python -m nasa_niswan_b200.build --knobs
export NINT_LIB=${NINT_LIB:-$(cd "$(dirname "$0")/.." && pwd)/nasa_niswan_b200/libnint_knobs.so}
for f in ${FLAGS:-0 1 2 3}; do
  echo "##### NINT_DEBUG_FLAGS=$f (1: no epilogue math/stores, 2: no MMA issue, 4: wgrad MMA stream twice, 8: timeline trace, 32: single MMA issuer)"
  NINT_DEBUG_FLAGS=$f python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d['kernels'].items(): print('  ', k, v['avg_launch_us'])
"
done
