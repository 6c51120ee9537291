#!/bin/bash
# N-GPU pass: NCCL test + bench at --gpus N.  Usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N>
TAG=${1:-m}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt
timeout 600 python -m pytest tests/test_parallel_nccl.py -m gpu -x -q > $OUT/${TAG}_pytest_nccl.log 2>&1; echo "pytest_nccl=$?"
tail -5 $OUT/${TAG}_pytest_nccl.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 100 --warmup 10 --skip-cpu-baseline > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench=$?"
tail -c 1500 $OUT/${TAG}_bench_n$N.json
tail -5 $OUT/${TAG}_bench_n$N.err
