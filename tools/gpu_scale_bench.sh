#!/bin/bash
# bench.py at N GPUs (the driver's scaling run).  Usage (under gpurun --gpus N): bash tools/gpu_scale_bench.sh <tag> <N>
TAG=${1:-s}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514"
timeout 600 $RUN bench.py --impl reference --gpus $N --steps 3 --warmup 3 > $OUT/${TAG}_bench_ref_n$N.json 2> $OUT/${TAG}_bench_ref_n$N.err; echo "bench_ref=$?"
timeout 600 $RUN bench.py --gpus $N --steps 300 --warmup 10 --skip-cpu-baseline > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench=$?"
python - <<PY
import json
d=json.loads(open("$OUT/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
print("n", d["n_gpus"], "value", d["value"], "e2e", d["e2e"]["value"], "stateless", d["e2e"]["stateless"]["value"], "gemm frac", d["roofline"]["frac"])
r=open("$OUT/${TAG}_bench_ref_n$N.json").read().strip().splitlines()
print("ref lines", len(r), r[-1][:120])
PY
