#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r2i_pytest.log 2>&1; echo "pytest=$?"; tail -8 $OUT/r2i_pytest.log
run() { tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline --skip-e2e > $OUT/r2i_$tag.json 2> $OUT/r2i_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/r2i_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "value %.4g ms %.4f" % (d["value"], d["ms_per_step"]), d.get("kernels_ms_per_step"))
except Exception as e:
    print("$tag failed", e); print(open("$OUT/r2i_$tag.err").read()[-1500:])
PY
}
run order A=1
run plain SIMSTEP_FINAL_ORDER=0
run order2 A=1
