#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/x_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/x_tests.log
tail -4 gpurun_out/x_tests.log
