#!/bin/bash
# configs[3]/[4] on N GPUs.  Usage (under gpurun [--gpus N]): bash tools/gpu_configs.sh <tag> <N>
TAG=${1:-c}; N=${2:-1}
OUT=gpurun_out; mkdir -p $OUT
if [ "$N" = "1" ]; then
  python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "scale_config" > $OUT/${TAG}_pytest_scale.log 2>&1; echo "pytest_scale=$?"; tail -3 $OUT/${TAG}_pytest_scale.log
  RUN="python"
else
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512"
fi
timeout 900 $RUN tools/bench_configs.py --config 4 > $OUT/${TAG}_config4_n$N.json 2> $OUT/${TAG}_config4_n$N.err; echo "config4=$?"; tail -c 1800 $OUT/${TAG}_config4_n$N.json; tail -3 $OUT/${TAG}_config4_n$N.err
timeout 900 $RUN tools/bench_configs.py --config 5 > $OUT/${TAG}_config5_n$N.json 2> $OUT/${TAG}_config5_n$N.err; echo "config5=$?"; tail -c 1800 $OUT/${TAG}_config5_n$N.json; tail -3 $OUT/${TAG}_config5_n$N.err
