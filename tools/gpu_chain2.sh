#!/bin/bash
# full GPU suite with the column-fused forward as the default, then an ncu capture of the chain kernel
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/chain2_pytest.log 2>&1; echo "pytest=$?"; tail -5 $OUT/chain2_pytest.log
python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-sustained --skip-extras > $OUT/chain2_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/chain2_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:ensemble_chain -s 4 -c 1 -o $OUT/chain2_prof \
  python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline --skip-sustained --skip-extras > $OUT/chain2_ncu.log 2>&1; echo "ncu=$?"
