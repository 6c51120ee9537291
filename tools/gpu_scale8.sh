#!/bin/bash
# 8-GPU pass: bench line + configs[3]/[4].  Usage (under gpurun --gpus 8): bash tools/gpu_scale8.sh <tag> [N]
TAG=${1:-s8}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $RUN bench.py --gpus $N --steps 300 --warmup 10 --skip-cpu-baseline > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench=$?"
tail -c 600 $OUT/${TAG}_bench_n$N.json | head -c 300; echo
timeout 600 $RUN tools/bench_configs.py --config 4 > $OUT/${TAG}_config4_n$N.json 2> $OUT/${TAG}_config4_n$N.err; echo "config4=$?"
head -c 250 $OUT/${TAG}_config4_n$N.json; echo
timeout 600 $RUN tools/bench_configs.py --config 5 > $OUT/${TAG}_config5_n$N.json 2> $OUT/${TAG}_config5_n$N.err; echo "config5=$?"
head -c 250 $OUT/${TAG}_config5_n$N.json; echo
