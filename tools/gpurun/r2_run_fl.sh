#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "time_fused" 2>&1 | tail -3
for f in 0 1 2 3 0 3; do
  NINT_FUSE_STEPS=$f timeout 300 python bench.py --no-extras --steps 20 --warmup 5 > gpurun_out/fl_bench_$f.json 2> gpurun_out/fl_err.log
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/fl_bench_$f.json") if l.startswith("{")][-1])
k = d["kernels"]
print("fuse=$f", d["value"], d["ms_per_step"], {n: (v["avg_launch_us"], v["kernel_launches_per_step"]) for n, v in k.items()}, d["gpu_launches"])
PY
done
for f in 0 1 2 3; do
  echo "shipped fuse=$f $(NINT_FUSE_STEPS=$f timeout 300 python tools/step_time.py --steps 30 --shipped 2>&1 | tail -1)"
done
