#!/bin/bash
# Re-capture the launch list and the K1 kernel of the bench command (after a K1 change).  Usage: bash scripts/gpu_profile_k1.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --val 1024"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:aggregate -c 4 -o gpurun_out/${TAG}_aggregate -f $CMD > gpurun_out/${TAG}_ncu_agg.log 2>&1
echo "aggregate full rc=$?"
