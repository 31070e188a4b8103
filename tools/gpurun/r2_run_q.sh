#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/q_plain.log 2>&1 || { tail -5 gpurun_out/q_plain.log; exit 1; }
ncu --set full --clock-control none -k 'regex:head_fwd_kernel|head_bwd_kernel|loss_mse_l1_kernel|adam_dev_kernel|unpack_wgrad_kernel|pack_cl_kernel|pack_w_fwd_kernel|pack_w_bwd_kernel' \
    -s 24 -c 8 -o gpurun_out/prof_small_r2c -f $CMD > gpurun_out/q_ncu_small.log 2>&1
echo "small-kernel capture rc=$?"
