#!/bin/bash
# Re-capture the launch list and the K1, LayerNorm and attention kernels of the bench command (after changing them).  Usage: bash scripts/gpu_profile_k1.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --val 1024"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:aggregate -c 4 -o gpurun_out/${TAG}_aggregate -f $CMD > gpurun_out/${TAG}_ncu_agg.log 2>&1
echo "aggregate full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:layernorm -s 4 -c 1 -o gpurun_out/${TAG}_layernorm -f $CMD > gpurun_out/${TAG}_ncu_ln.log 2>&1
echo "layernorm full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 4 -c 1 -o gpurun_out/${TAG}_attention -f $CMD > gpurun_out/${TAG}_ncu_att.log 2>&1
echo "attention full rc=$?"
