#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
export SIMSTEP_FINAL_FUSED=0
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_simenv_ref_gpu.py tests/test_rollout_gpu.py tests/test_host_api_gpu.py tests/test_train_gpu.py -m gpu -q > $OUT/r2g_pytest.log 2>&1; echo "pytest=$?"; tail -12 $OUT/r2g_pytest.log
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline --skip-e2e > $OUT/r2g_$tag.json 2> $OUT/r2g_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/r2g_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "value %.4g ms %.4f" % (d["value"], d["ms_per_step"]), d.get("kernels_ms_per_step"))
except Exception as e:
    print("$tag failed", e); print(open("$OUT/r2g_$tag.err").read()[-1500:])
PY
}
run wide A=1
run narrow SIMSTEP_GEMM_WIDE_EPI=0
