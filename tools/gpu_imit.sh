set -u
python -m pytest tests/test_imitation_gpu.py -x -q > gpurun_out/r01e_pytest_imit.log 2>&1; echo "pytest=$?"
python tools/bench_imitation.py > gpurun_out/r01e_imit.json 2> gpurun_out/r01e_imit.err; echo "imit=$?"
python tools/bench_imitation.py --origin --terms >> gpurun_out/r01e_imit.json 2>> gpurun_out/r01e_imit.err
python tools/bench_imitation.py --iters 3 --warmup 1 > gpurun_out/r01e_plain_imit.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:imitation_reward -s 1 -c 1 \
  -o gpurun_out/r01e_prof_imit python tools/bench_imitation.py --iters 3 --warmup 1 > gpurun_out/r01e_ncu_imit.log 2>&1; echo "ncu_imit=$?"
