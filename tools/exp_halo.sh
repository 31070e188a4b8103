#!/bin/bash
# hardware experiment: halo-variant correctness under descriptor base-offset policies and cluster sizes
for bo in 0 1; do
  echo "##### NINT_CLUSTER=1 NINT_BASE_OFFSET=$bo"
  NINT_CLUSTER=1 NINT_BASE_OFFSET=$bo timeout 120 python tools/gpu_probe.py raw_bf16_small raw_bf16 raw_tf32 2>&1 | grep -v Warn | tail -12
done
for cl in 2 4; do
  echo "##### NINT_CLUSTER=$cl NINT_BASE_OFFSET=0"
  NINT_CLUSTER=$cl NINT_BASE_OFFSET=0 timeout 120 python tools/gpu_probe.py raw_bf16_small raw_bf16 2>&1 | grep -v Warn | tail -12
done
