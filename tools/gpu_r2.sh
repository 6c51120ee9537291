#!/bin/bash
# Round-2 development pass: parity tests, then the bench line with the fused final layer on and off.
# Usage (under gpurun): bash tools/gpu_r2.sh <tag> [pytest targets...]
TAG=${1:-r2}; shift
OUT=gpurun_out; mkdir -p $OUT
TARGETS=${@:-tests}
timeout 1500 python -m pytest $TARGETS -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?"
tail -25 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline --skip-e2e > $OUT/${TAG}_bench_fused.json 2> $OUT/${TAG}_bench_fused.err; echo "bench_fused=$?"
SIMSTEP_FINAL_FUSED=0 timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline --skip-e2e > $OUT/${TAG}_bench_legacy.json 2> $OUT/${TAG}_bench_legacy.err; echo "bench_legacy=$?"
python - <<PY
import json
for k in ("fused", "legacy"):
    try:
        d = json.loads(open("$OUT/${TAG}_bench_%s.json" % k).read().strip().splitlines()[-1])
        print(k, "value %.4g ms %.4f" % (d["value"], d["ms_per_step"]), d.get("kernels_ms_per_step"))
    except Exception as e:
        print(k, "no line:", e)
        print(open("$OUT/${TAG}_bench_%s.err" % k).read()[-1500:])
PY
