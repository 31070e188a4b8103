#!/bin/bash
# same-box A/B of two builds of the library (NINT_LIB): libnint_prev.so vs libnint.so
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
k = d["kernels"]
print(sys.argv[1], d["value"], d["ms_per_step"], {n: v["avg_launch_us"] for n, v in k.items()})
PY
}
for rep in 1 2 3; do
  NINT_LIB=$PWD/nasa_niswan_b200/libnint_prev.so timeout 300 python bench.py --no-extras --steps 20 --warmup 5 > gpurun_out/wg_prev.json 2> gpurun_out/wg_prev.err; show prev gpurun_out/wg_prev.json
  timeout 300 python bench.py --no-extras --steps 20 --warmup 5 > gpurun_out/wg_new.json 2> gpurun_out/wg_new.err; show new gpurun_out/wg_new.json
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -m gpu -q -x 2>&1 | tail -2
