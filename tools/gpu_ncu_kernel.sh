#!/bin/bash
# ncu --set full capture of one kernel of the bench step.  Usage: bash tools/gpu_ncu_kernel.sh <tag> <kernel regex> [skip] [count]
TAG=$1; RE=$2; SKIP=${3:-5}; CNT=${4:-1}
OUT=gpurun_out; mkdir -p $OUT
python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o $OUT/${TAG}_prof \
  python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1; echo "ncu=$?"
