#!/bin/bash
# timing matrix over SIMSTEP_FINAL_DEBUG switches (results are wrong when a switch is set; timing only)
OUT=gpurun_out; mkdir -p $OUT
for D in "$@"; do
  SIMSTEP_FINAL_DEBUG=$D timeout 200 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-e2e > $OUT/dbg_$D.json 2> $OUT/dbg_$D.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/dbg_$D.json").read().strip().splitlines()[-1])
    print("debug=$D ms %.4f gemm %.4f" % (d["ms_per_step"], d["kernels_ms_per_step"]["ensemble_gemm"]))
except Exception as e:
    print("debug=$D failed", e); print(open("$OUT/dbg_$D.err").read()[-800:])
PY
done
