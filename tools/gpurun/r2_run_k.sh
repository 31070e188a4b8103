#!/bin/bash
# 8-GPU pass: scaling with the fused NVLink step tail (and NCCL for comparison), e2e variants at 8 ranks
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/k_topo.txt 2>&1
lscpu | grep -E "^CPU\(s\)|NUMA|Model name|Socket" >> gpurun_out/k_topo.txt
run() {  # $1 = fused, rest = bench args
  f=$1; shift
  NINT_DP_FUSED=$f timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
     --master-port $((29500 + RANDOM % 500)) bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline "$@" 2>> gpurun_out/k_bench.err
}
run 1 > gpurun_out/k_bench8_fused.json
run 0 --no-extras > gpurun_out/k_bench8_nccl.json
run 1 --no-extras > gpurun_out/k_bench8_fused_b.json
run 0 --no-extras > gpurun_out/k_bench8_nccl_b.json
python - <<'PY'
import json
for n in ("k_bench8_fused", "k_bench8_nccl", "k_bench8_fused_b", "k_bench8_nccl_b"):
    try:
        e = json.loads(open(f"gpurun_out/{n}.json").readline())
        print(n, e["value"], e["ms_per_step"], e["gate_conv_fwd_bwd"]["other_ms_per_step"], e["e2e"]["value"],
              {k: v["value"] for k, v in e.get("e2e_variants", {}).items()}, e.get("sustained"), e.get("frame_bank_resident"))
    except Exception as ex:
        print(n, "failed", ex)
PY
tail -5 gpurun_out/k_bench.err
