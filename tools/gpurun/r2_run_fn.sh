#!/bin/bash
# final-state regression: full GPU suite, default bench, ncu launch list, one full capture of a BPTT launch (t = 5)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/fn_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/fn_tests.log
tail -4 gpurun_out/fn_tests.log
timeout 900 python bench.py > gpurun_out/fn_bench.json 2> gpurun_out/fn_bench.err
echo "bench rc=$?"; tail -1 gpurun_out/fn_bench.json | cut -c1-600
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$CMD > gpurun_out/fn_plain.log 2>&1 || { echo "plain failed"; exit 0; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 75 -c 60 --csv --log-file gpurun_out/fn_launches.csv $CMD > gpurun_out/fn_ncu_list.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:conv_halo_kernel' -s 90 -c 1 -o gpurun_out/fn_prof_bwd -f $CMD > gpurun_out/fn_ncu_bwd.log 2>&1
echo "bwd capture rc=$?"; tail -2 gpurun_out/fn_ncu_bwd.log | cut -c1-200
for f in 0 2; do echo "det fuse=$f $(NINT_FUSE_STEPS=$f timeout 300 python tools/fused_stress.py 300 --det 2>&1 | grep 'steps ok\|code=[1-9]\|Error')"; done
NINT_FUSE_STEPS=3 timeout 300 python tools/fused_stress.py 2000 2>&1 | grep 'steps ok\|code=[1-9]\|Error' 
