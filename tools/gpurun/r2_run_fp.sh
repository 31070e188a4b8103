#!/bin/bash
# same-box A/B: tree of the commit before the time-fused work (ab_old/) against the current tree
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
k = d["kernels"]
print(sys.argv[1], d["value"], d["ms_per_step"], {n: v["avg_launch_us"] for n, v in k.items()}, d["gate_conv_fwd_bwd"]["other_ms_per_step"], d["clocks"].get("power_w"))
PY
}
for rep in 1 2 3; do
  (cd ab_old && timeout 300 python bench.py --no-extras --steps 20 --warmup 5 > ../gpurun_out/fp_old.json 2> ../gpurun_out/fp_old.err); show old gpurun_out/fp_old.json
  timeout 300 python bench.py --no-extras --steps 20 --warmup 5 > gpurun_out/fp_new.json 2> gpurun_out/fp_new.err; show new gpurun_out/fp_new.json
  NINT_FUSE_STEPS=0 timeout 300 python bench.py --no-extras --steps 20 --warmup 5 > gpurun_out/fp_new0.json 2> gpurun_out/fp_new0.err; show new_fuse0 gpurun_out/fp_new0.json
done
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,temperature.gpu --format=csv
