#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (both arms), imitation microbench, ncu launch list + full captures.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest=$?"
python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke=$?"
python bench.py --impl reference --steps 5 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "bench_ref=$?"
python bench.py --steps 300 --warmup 10 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench=$?"
python tools/bench_imitation.py > $OUT/${TAG}_imit.json 2> $OUT/${TAG}_imit.err; echo "imit=$?"
python tools/bench_imitation.py --origin --terms --skip-cpu >> $OUT/${TAG}_imit.json 2>> $OUT/${TAG}_imit.err
python tools/bench_rollout.py > $OUT/${TAG}_rollout.json 2> $OUT/${TAG}_rollout.err; echo "rollout=$?"
python tools/bench_mlpcost.py > $OUT/${TAG}_mlpcost.json 2> $OUT/${TAG}_mlpcost.err; echo "mlpcost=$?"
if [ "${NCU:-1}" = "1" ]; then
python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1; echo "ncu_launches=$?"
ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05|post_step|prep_input|rff_pack|cost_combine' -s 30 -c 10 \
  -o $OUT/${TAG}_prof_step python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline > $OUT/${TAG}_ncu_step.log 2>&1; echo "ncu_step=$?"
python tools/bench_imitation.py --iters 3 --warmup 1 > $OUT/${TAG}_plain_imit.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:imitation_reward -s 1 -c 1 \
  -o $OUT/${TAG}_prof_imit python tools/bench_imitation.py --iters 3 --warmup 1 > $OUT/${TAG}_ncu_imit.log 2>&1; echo "ncu_imit=$?"
fi
