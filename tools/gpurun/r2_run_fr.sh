#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "time_fused" 2>&1 | tail -2
bash tools/gpurun/r2_run_fp.sh
