#!/bin/bash
# Round 2, on the GPU box (under gpurun): plain run first, then the ncu launch list, one full capture per tensor kernel
# and full captures of the memory-bound kernels.  Usage: bash tools/ncu_capture_r2.sh <tag>
set -u
TAG=${1:-r2}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log | cut -c1-400
# launch list of the two timed steps (3 warm-up steps of ~35 launches skipped)
ncu --metrics gpu__time_duration.sum --clock-control none -s 105 -c 80 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
# full captures: forward conv at t>=1 (launch 73 = 3 warm-up steps * 24 conv launches + t=1), a backward conv, wgrad
ncu --set full --clock-control none --import-source on -k 'regex:conv_halo_kernel' -s 73 -c 1 \
    -o gpurun_out/prof_fwd_$TAG -f $CMD > gpurun_out/ncu_fwd_$TAG.log 2>&1
echo "fwd capture rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:conv_halo_kernel' -s 86 -c 1 \
    -o gpurun_out/prof_bwd_$TAG -f $CMD > gpurun_out/ncu_bwd_$TAG.log 2>&1
echo "bwd capture rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:^wgrad_(pair_)?kernel' -s 3 -c 1 \
    -o gpurun_out/prof_wgrad_$TAG -f $CMD > gpurun_out/ncu_wgrad_$TAG.log 2>&1
echo "wgrad capture rc=$?"
# memory-bound kernels of the step (4th step: one launch each) and of the preprocessing path
ncu --set full --clock-control none -k 'regex:head_fwd_kernel|head_bwd_kernel|loss_mse_l1_kernel|adam_dev_kernel|unpack_wgrad_kernel|pack_cl_kernel|pack_w_fwd_kernel|pack_w_bwd_kernel' \
    -s 24 -c 8 -o gpurun_out/prof_small_$TAG -f $CMD > gpurun_out/ncu_small_$TAG.log 2>&1
echo "small-kernel capture rc=$?"
python tools/pointwise_bench.py --once > gpurun_out/pointwise_plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none -k 'regex:fuse_kernel|pack_cl_kernel' -c 6 \
    -o gpurun_out/prof_prep_$TAG -f python tools/pointwise_bench.py --once > gpurun_out/ncu_prep_$TAG.log 2>&1
echo "preprocessing capture rc=$?"
ls -la gpurun_out/ | grep $TAG
