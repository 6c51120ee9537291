#!/bin/bash
# column-fused forward (gemm_chain.cuh) against one launch per layer: bitwise comparison of the step's outputs, then a
# short bench per mode.  Usage: bash tools/gpu_chain.sh [tag]
TAG=${1:-chain}
OUT=gpurun_out; mkdir -p $OUT
SIMSTEP_CHAIN=0 timeout 300 python tools/final_fused_check.py /tmp/c0.npz || echo "per-layer run failed"
SIMSTEP_CHAIN=1 SIMSTEP_CHAIN_SLOT=0 timeout 300 python tools/final_fused_check.py /tmp/c1.npz || echo "chain run failed"
SIMSTEP_CHAIN=1 SIMSTEP_CHAIN_SLOT=1 timeout 300 python tools/final_fused_check.py /tmp/c2.npz || echo "chain slot run failed"
python - <<'PY'
import numpy as np
a = np.load("/tmp/c0.npz")
for name in ("/tmp/c1.npz", "/tmp/c2.npz"):
    try:
        b = np.load(name)
    except Exception as e:
        print(name, "missing", e); continue
    bad = [k for k in a.files if not np.array_equal(a[k], b[k], equal_nan=True)]
    print(name, "bitwise identical" if not bad else "DIFFERS in %s" % bad[:8])
    for k in bad[:4]:
        print("  ", k, float(np.nanmax(np.abs(a[k].astype(np.float64) - b[k].astype(np.float64)))))
PY
for mode in "SIMSTEP_CHAIN=0" "SIMSTEP_CHAIN_SLOT=0" "SIMSTEP_CHAIN_SLOT=1"; do
for i in 1 2; do
env $mode timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline --skip-e2e --skip-sustained --skip-extras > $OUT/${TAG}_${mode}_$i.json 2> $OUT/${TAG}_${mode}_$i.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_${mode}_$i.json").read().strip().splitlines()[-1])
    print("$mode run $i", "value %.4g ms %.4f" % (d["value"], d["ms_per_step"]), d.get("kernels_ms_per_step"), d.get("clocks"))
except Exception as e:
    print("$mode failed", e); print(open("$OUT/${TAG}_${mode}_$i.err").read()[-1500:])
PY
done
done
