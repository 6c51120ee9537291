#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r2j_pytest.log 2>&1; echo "pytest=$?"; tail -12 $OUT/r2j_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/r2j_bench.json 2> $OUT/r2j_bench.err; echo "bench=$?"
tail -5 $OUT/r2j_bench.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/r2j_bench.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","e2e","sustained","roofline","roofline_hbm","kernels_ms_per_step","plugin_single_env","e2e_device_policy","cpu_baseline","gpu_launches","clocks"):
        print(k, json.dumps(d.get(k))[:700])
except Exception as e:
    print("no line", e)
PY
