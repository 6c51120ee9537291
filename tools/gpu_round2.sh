#!/bin/bash
# Round-2 GPU pass: full parity suite, smoke, both bench arms, configs 4/5 at N=1, ncu launch list + full captures.
# Usage (under gpurun): bash tools/gpu_round2.sh <tag>
set -u
TAG=${1:-r02a}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest=$?"; tail -3 $OUT/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke=$?"; tail -1 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench_ref=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?"
timeout 900 python bench.py --config 4 --skip-cpu-baseline > $OUT/${TAG}_bench_c4.json 2> $OUT/${TAG}_bench_c4.err; echo "bench_c4=$?"
timeout 900 python bench.py --config 5 --skip-cpu-baseline > $OUT/${TAG}_bench_c5.json 2> $OUT/${TAG}_bench_c5.err; echo "bench_c5=$?"
timeout 300 python tools/bench_imitation.py > $OUT/${TAG}_imit.json 2> $OUT/${TAG}_imit.err; echo "imit=$?"
CMD="python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-sustained"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1; echo "ncu_launches=$?"
ncu --set full --clock-control none --import-source on -k regex:'ensemble_chain|gemm_tcgen05|post_step|prep_input' -s ${NCU_SKIP:-12} -c ${NCU_COUNT:-4} \
  -o $OUT/${TAG}_prof_step $CMD > $OUT/${TAG}_ncu_step.log 2>&1; echo "ncu_step=$?"
python tools/bench_imitation.py --iters 3 --warmup 1 > $OUT/${TAG}_plain_imit.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:imitation_reward -s 1 -c 1 \
  -o $OUT/${TAG}_prof_imit python tools/bench_imitation.py --iters 3 --warmup 1 > $OUT/${TAG}_ncu_imit.log 2>&1; echo "ncu_imit=$?"
