#!/bin/bash
# last check of the round: whole GPU suite and a short bench with programmatic dependent launch off (the new default)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 100 python bench.py --no-extras --steps 20 --warmup 5 > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err
echo "bench rc=$?"; grep '^{' gpurun_out/last_bench.json | tail -1 | cut -c1-160
