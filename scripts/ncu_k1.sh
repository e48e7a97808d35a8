#!/bin/bash
# ncu --set full of one K1 shape.  Usage: bash scripts/ncu_k1.sh <tag> N C [f16|f32]
TAG=$1; shift
mkdir -p gpurun_out
python scripts/k1_probe.py "$@" > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:aggregate -s 2 -c 1 -o gpurun_out/${TAG}_k1 -f python scripts/k1_probe.py "$@" > gpurun_out/${TAG}_ncu.log 2>&1
cat gpurun_out/${TAG}_plain.log
