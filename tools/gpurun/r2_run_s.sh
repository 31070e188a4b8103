#!/bin/bash
# BASELINE cfg 3 (global batch 256) at N ranks + the weak-scaling line; usage: bash tools/gpurun/r2_run_s.sh N
cd "$(dirname "$0")/../.."
N=$1
mkdir -p gpurun_out
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
     --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline "$@" 2>> gpurun_out/s_bench$N.err | grep "^{"
}
run --global-batch 256 --no-extras > gpurun_out/s_cfg3_${N}gpu.json
run > gpurun_out/s_weak_${N}gpu.json
python - <<PY
import json
for n in ("s_cfg3_${N}gpu", "s_weak_${N}gpu"):
    e = json.load(open(f"gpurun_out/{n}.json"))
    print(n, e["value"], e["ms_per_step"], e["scaling"], e["config"]["workload"][-40:], e["gate_conv_fwd_bwd"]["other_ms_per_step"], e["e2e"]["value"],
          {k: v["value"] for k, v in e.get("e2e_variants", {}).items()})
PY
