#!/bin/bash
# the env step's tail inside the chain kernel (SIMSTEP_CHAIN_TAIL=1) against the default (post-step kernel)
OUT=gpurun_out; mkdir -p $OUT
SIMSTEP_CHAIN_TAIL=0 timeout 300 python tools/final_fused_check.py /tmp/t0.npz || echo "default run failed"
SIMSTEP_CHAIN_TAIL=1 timeout 300 python tools/final_fused_check.py /tmp/t1.npz 2>&1 | tail -5
python - <<'PY'
import numpy as np
a, b = np.load("/tmp/t0.npz"), np.load("/tmp/t1.npz")
for k in a.files:
    x, y = a[k], b[k]
    if k.endswith(("_next", "_done", "_steps")):
        if not np.array_equal(x, y, equal_nan=True):
            print("DIFF (must be bitwise)", k, int((x != y).sum()))
    else:
        scale = max(float(np.nanmax(np.abs(x))), 1e-12)
        d = float(np.nanmax(np.abs(x.astype(np.float64) - y.astype(np.float64))))
        if d > 2e-6 * scale:
            print("DIFF", k, d, scale)
print("compared", len(a.files), "arrays")
PY
bash tools/gpu_ab2.sh tail SIMSTEP_CHAIN_TAIL=0 SIMSTEP_CHAIN_TAIL=1
