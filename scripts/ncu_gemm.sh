#!/bin/bash
# ncu --set full of one GEMM probe shape.  Usage: bash scripts/ncu_gemm.sh <tag> <shape> [env...]
TAG=$1; SHAPE=$2
python scripts/gemm_probe.py $SHAPE > gpurun_out/${TAG}_${SHAPE}_plainrun.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 4 -c 1 -o gpurun_out/${TAG}_gemm_${SHAPE} -f python scripts/gemm_probe.py $SHAPE > gpurun_out/${TAG}_ncu_${SHAPE}.log 2>&1
cat gpurun_out/${TAG}_${SHAPE}_plainrun.log
