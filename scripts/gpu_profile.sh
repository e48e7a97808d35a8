#!/bin/bash
# ncu evidence for the bench command (reduced validation set so the launch list stays short).
# Usage: bash scripts/gpu_profile.sh <tag>      -> gpurun_out/<tag>_*.{csv,ncu-rep,log}
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --val 1024"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
# one encoder layer's four GEMMs (QKV, out-proj, MLP-up, MLP-down): skip the patch-embedding GEMM and 7 layers
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 29 -c 4 -o gpurun_out/${TAG}_gemm -f $CMD > gpurun_out/${TAG}_ncu_gemm.log 2>&1
echo "gemm full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:aggregate -c 4 -o gpurun_out/${TAG}_aggregate -f $CMD > gpurun_out/${TAG}_ncu_agg.log 2>&1
echo "aggregate full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_tc_kernel -s 4 -c 1 -o gpurun_out/${TAG}_attention -f $CMD > gpurun_out/${TAG}_ncu_att.log 2>&1
echo "attention full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 4 -c 1 -o gpurun_out/${TAG}_layernorm -f $CMD > gpurun_out/${TAG}_ncu_ln.log 2>&1
echo "layernorm full rc=$?"
# the split-precision mode (f16x3): one layer's GEMMs, the split pass and the split attention
CMD3="$CMD --precision f16x3"
$CMD3 > gpurun_out/${TAG}_f16x3_plain.log 2>&1 || { echo "f16x3 plain run failed"; tail -5 gpurun_out/${TAG}_f16x3_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_f16x3_plain.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 29 -c 4 -o gpurun_out/${TAG}_f16x3_gemm -f $CMD3 > gpurun_out/${TAG}_ncu_f16x3_gemm.log 2>&1
echo "f16x3 gemm full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_split_kernel -s 4 -c 1 -o gpurun_out/${TAG}_f16x3_attention -f $CMD3 > gpurun_out/${TAG}_ncu_f16x3_att.log 2>&1
echo "f16x3 attention full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:split_f16_kernel -s 20 -c 2 -o gpurun_out/${TAG}_f16x3_split -f $CMD3 > gpurun_out/${TAG}_ncu_f16x3_split.log 2>&1
echo "f16x3 split full rc=$?"
ls -la gpurun_out/ | grep ${TAG}
