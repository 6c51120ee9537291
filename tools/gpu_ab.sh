#!/bin/bash
# quick A/B: selected parity tests + short bench (prints per-kernel times).  Usage: bash tools/gpu_ab.sh <tag> [env...]
TAG=${1:-ab}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_simenv_ref_gpu.py tests/test_full_size_parity_gpu.py -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do
env "$@" timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline --skip-e2e --skip-sustained > $OUT/${TAG}_$i.json 2> $OUT/${TAG}_$i.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_$i.json").read().strip().splitlines()[-1])
    print("$TAG run $i", "value %.4g ms %.4f" % (d["value"], d["ms_per_step"]), d.get("kernels_ms_per_step"))
except Exception as e:
    print("$TAG failed", e); print(open("$OUT/${TAG}_$i.err").read()[-1500:])
PY
done
