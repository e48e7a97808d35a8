#!/bin/bash
# last session of a round: everything green with the final build, the headline lines, the GEMM capture refreshed
bash scripts/gpu_round.sh
TAG=${1:-r1z}
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --val 1024"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 29 -c 4 -o gpurun_out/${TAG}_gemm -f $CMD > gpurun_out/${TAG}_ncu_gemm.log 2>&1
echo "gemm full rc=$?"
