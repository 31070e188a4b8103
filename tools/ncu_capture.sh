#!/bin/bash
# Run on the GPU box (under gpurun): plain run first, then the ncu launch list and one full capture
# per tensor kernel.  Usage: bash tools/ncu_capture.sh <tag>
set -u
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log
# launch list of the two timed steps (3 warm-up steps = 3*~33 launches skipped)
if [ -z "${SKIP_LIST:-}" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 130 --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
fi
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
ls -la gpurun_out/
