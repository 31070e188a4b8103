#!/bin/bash
# intermittent launch failure of the one-launch-per-step schedule on the reference's recipe under programmatic
# dependent launch: 12 runs of 150 deterministic training steps with PDL on (it failed 4 of 8 before the fix)
cd "$(dirname "$0")/../.."
ok1=0; bad1=0
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(NINT_FUSE_STEPS=0 timeout 300 python tools/fused_stress.py 150 --shipped --det 2>&1 | grep -c 'steps ok')
  if [ "$out" = "1" ]; then ok1=$((ok1+1)); else bad1=$((bad1+1)); fi
done
echo "pdl=1 with the second rendezvous: ok $ok1 bad $bad1"
