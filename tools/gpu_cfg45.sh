#!/bin/bash
# configs[3] and configs[4] with the column-fused forward against one launch per layer
OUT=gpurun_out; mkdir -p $OUT
for cfg in 5 4; do
for m in 0 2; do
SIMSTEP_CHAIN=$m timeout 600 python bench.py --config $cfg --steps 10 --warmup 3 --skip-cpu-baseline --skip-e2e --skip-sustained --skip-extras > $OUT/cfg${cfg}_m${m}.json 2> $OUT/cfg${cfg}_m${m}.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/cfg${cfg}_m${m}.json").read().strip().splitlines()[-1])
    print("config $cfg chain mode $m", "value %.4g ms %.4f" % (d["value"], d["ms_per_step"]), {k: round(v, 4) for k, v in (d.get("kernels_ms_per_step") or {}).items()})
except Exception as e:
    print("config $cfg mode $m failed", e); print(open("$OUT/cfg${cfg}_m${m}.err").read()[-1500:])
PY
done
done
