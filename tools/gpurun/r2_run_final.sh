#!/bin/bash
# last check of the committed tree: smoke(), the whole GPU suite, the default bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
echo "bench rc=$?"; grep '^{' gpurun_out/final_bench.json | tail -1 | cut -c1-200
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | grep '^{' | cut -c1-250
