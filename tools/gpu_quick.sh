#!/bin/bash
# parity tests + bench line (no ncu).  Usage: bash tools/gpu_quick.sh <tag> [bench args]
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest=$?"
python bench.py --steps 300 --warmup 10 --skip-cpu-baseline "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?"
