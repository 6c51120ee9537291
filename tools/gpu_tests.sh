#!/bin/bash
# selected GPU tests only.  Usage: bash tools/gpu_tests.sh <tag> <pytest args...>
TAG=${1:-t}; shift
OUT=gpurun_out; mkdir -p $OUT
python -m pytest "$@" -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest=$?"
tail -30 $OUT/${TAG}_pytest.log
