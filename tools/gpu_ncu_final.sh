#!/bin/bash
# ncu --set full on the fused final kernel (and the legacy final GEMM for comparison)
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2}
CMD="python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline"
$CMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:final_fused -s 4 -c 2 -o $OUT/${TAG}_prof_final $CMD > $OUT/${TAG}_ncu_final.log 2>&1; echo "ncu_final=$?"
SIMSTEP_FINAL_FUSED=0 $CMD > $OUT/${TAG}_plain_legacy.log 2>&1 &&
SIMSTEP_FINAL_FUSED=0 ncu --set full --clock-control none --import-source on -k regex:'gemm_tcgen05_kernel.*Li1ELi2' -s 4 -c 2 -o $OUT/${TAG}_prof_legacy $CMD > $OUT/${TAG}_ncu_legacy.log 2>&1; echo "ncu_legacy=$?"
tail -5 $OUT/${TAG}_ncu_final.log $OUT/${TAG}_ncu_legacy.log
