#!/bin/bash
# ncu capture of the chain kernel in a given mode.  Usage: bash tools/gpu_chain_ncu.sh <mode> <tag>
M=$1; TAG=$2
OUT=gpurun_out; mkdir -p $OUT
export SIMSTEP_CHAIN=$M SIMSTEP_CHAIN_DEBUG=1
python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-sustained --skip-extras > $OUT/${TAG}_plain.log 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
grep "simstep:" $OUT/${TAG}_plain.err
ncu --set full --clock-control none --import-source on -k regex:ensemble_chain -s 4 -c 1 -o $OUT/${TAG}_prof \
  python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-sustained --skip-extras > $OUT/${TAG}_ncu.log 2>&1; echo "ncu=$?"
