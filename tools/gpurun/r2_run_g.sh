#!/bin/bash
# 2-GPU pass: fused NVLink step tail vs NCCL, model on a non-current device
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/g_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "fused_nvlink or non_current_device" > gpurun_out/g_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/g_tests.log
tail -30 gpurun_out/g_tests.log
for fused in 1 0 1 0; do
  NINT_DP_FUSED=$fused timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
     bench.py --gpus 2 --steps 40 --warmup 5 --no-extras --no-cpu-baseline >> gpurun_out/g_bench_fused$fused.json 2>> gpurun_out/g_bench.err
done
python - <<'PY'
import json
for f in (1, 0):
    for line in open(f"gpurun_out/g_bench_fused{f}.json"):
        try:
            e = json.loads(line)
        except ValueError:
            continue
        print("fused", f, e["value"], e["ms_per_step"], e["gate_conv_fwd_bwd"]["other_ms_per_step"], e["e2e"]["value"])
PY
tail -5 gpurun_out/g_bench.err
