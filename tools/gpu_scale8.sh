#!/bin/bash
# N-GPU pass: the bench line (configs[1]) + configs[3] / configs[4] through bench.py --config, plus the PCIe probe.
# Usage (under gpurun --gpus N): bash tools/gpu_scale8.sh <tag> [N]
TAG=${1:-s8}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt; lscpu > $OUT/${TAG}_lscpu.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $RUN bench.py --gpus $N --steps 100 --warmup 10 --skip-cpu-baseline > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench=$?"
timeout 600 $RUN bench.py --gpus $N --config 4 --skip-cpu-baseline > $OUT/${TAG}_bench_c4_n$N.json 2> $OUT/${TAG}_bench_c4_n$N.err; echo "config4=$?"
timeout 600 $RUN bench.py --gpus $N --config 5 --skip-cpu-baseline > $OUT/${TAG}_bench_c5_n$N.json 2> $OUT/${TAG}_bench_c5_n$N.err; echo "config5=$?"
python - <<PY
import json
for t in ("bench", "bench_c4", "bench_c5"):
    try:
        d = json.loads(open("$OUT/${TAG}_%s_n$N.json" % t).read().strip().splitlines()[-1])
        print(t, "value %.4g ms %.3f" % (d["value"], d["ms_per_step"]), "e2e %.4g" % d["e2e"]["value"], "parity", d.get("shard_parity"),
              "quantile", d.get("quantile", {}).get("device_ms_per_call"), "pcie", d["e2e"].get("pcie_d2h_probe_gbs_per_rank"))
    except Exception as e:
        print(t, "no line", e); print(open("$OUT/${TAG}_%s_n$N.err" % t).read()[-1200:])
PY
