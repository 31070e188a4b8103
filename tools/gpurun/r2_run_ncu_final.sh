#!/bin/bash
# full ncu captures of the forward conv and the wgrad kernel on the final build (the BPTT capture is in r2_run_fn.sh)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/nf_plain.log 2>&1 || { echo "plain failed"; exit 0; }
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:conv_halo_kernel' -s 73 -c 1 -o gpurun_out/nf_prof_fwd -f $CMD > gpurun_out/nf_ncu_fwd.log 2>&1
echo "fwd capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:^wgrad_(pair_)?kernel' -s 3 -c 1 -o gpurun_out/nf_prof_wgrad -f $CMD > gpurun_out/nf_ncu_wgrad.log 2>&1
echo "wgrad capture rc=$?"
