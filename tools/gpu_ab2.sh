#!/bin/bash
# A/B of env switches on the short bench.  Usage: bash tools/gpu_ab2.sh <tag> "<ENV=..>" "<ENV=..>" ...
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
k=0
for envs in "$@"; do
k=$((k+1))
for i in 1 2; do
env $envs timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline --skip-e2e --skip-sustained --skip-extras > $OUT/${TAG}_${k}_$i.json 2> $OUT/${TAG}_${k}_$i.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_${k}_$i.json").read().strip().splitlines()[-1])
    print("$envs run $i", "value %.4g ms %.4f" % (d["value"], d["ms_per_step"]), {k: round(v, 4) for k, v in d.get("kernels_ms_per_step").items()})
except Exception as e:
    print("$envs failed", e); print(open("$OUT/${TAG}_${k}_$i.err").read()[-1500:])
PY
done
done
