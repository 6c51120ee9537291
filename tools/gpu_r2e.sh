#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
export SIMSTEP_FINAL_FUSED=0
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_simenv_ref_gpu.py tests/test_rollout_gpu.py tests/test_host_api_gpu.py -m gpu -q > $OUT/r2e_pytest.log 2>&1; echo "pytest=$?"; tail -15 $OUT/r2e_pytest.log
for M in auto on off; do
timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline --skip-e2e --rff-split $M > $OUT/r2e_bench_$M.json 2> $OUT/r2e_bench_$M.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/r2e_bench_$M.json").read().strip().splitlines()[-1])
    print("$M", "value %.4g ms %.4f" % (d["value"], d["ms_per_step"]), d.get("kernels_ms_per_step"))
except Exception as e:
    print("$M failed", e); print(open("$OUT/r2e_bench_$M.err").read()[-1500:])
PY
done
SIMSTEP_RFF_FUSED_COMBINE=0 timeout 300 python bench.py --steps 50 --warmup 5 --skip-cpu-baseline --skip-e2e --rff-split on 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('unfused-combine on', d['ms_per_step'], d.get('kernels_ms_per_step'))"
grep "split decision" $OUT/r2e_bench_auto.err
