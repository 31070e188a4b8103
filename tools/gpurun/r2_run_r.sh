#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "golden or return_sequence or val_loop or sensitivity or rollout or native or determin or staged" > gpurun_out/r_tests.log 2>&1; tail -2 gpurun_out/r_tests.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/r_plain.log 2>&1 || { tail -5 gpurun_out/r_plain.log; exit 1; }
ncu --set full --clock-control none -k 'regex:head_fwd_kernel|head_bwd_kernel|loss_mse_l1_kernel|adam_dev_kernel|unpack_wgrad_kernel|pack_cl_kernel|pack_w_fwd_kernel|pack_w_bwd_kernel' \
    -s 24 -c 8 -o gpurun_out/prof_small_r2d -f $CMD > gpurun_out/r_ncu_small.log 2>&1
echo "small-kernel capture rc=$?"
