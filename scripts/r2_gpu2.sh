#!/bin/bash
# Round-2 GPU session 2: the default (gate-passing, f16c8) bench line at full size, the reference arm from the staged
# copy, the ncu launch list and --set full captures of one layer's kernels.
mkdir -p gpurun_out
TAG=r2
python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/${TAG}_bench_default.json
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "reference rc=$?"
cut -c1-300 gpurun_out/${TAG}_bench_reference.json; tail -3 gpurun_out/${TAG}_bench_reference.err
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-throughput-mode --val 1024"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 29 -c 4 -o gpurun_out/${TAG}_gemm -f $CMD > gpurun_out/${TAG}_ncu_gemm.log 2>&1
echo "gemm full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:aggregate -c 4 -o gpurun_out/${TAG}_aggregate -f $CMD > gpurun_out/${TAG}_ncu_agg.log 2>&1
echo "aggregate full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention -s 4 -c 1 -o gpurun_out/${TAG}_attention -f $CMD > gpurun_out/${TAG}_ncu_att.log 2>&1
echo "attention full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:layernorm -s 4 -c 1 -o gpurun_out/${TAG}_layernorm -f $CMD > gpurun_out/${TAG}_ncu_ln.log 2>&1
echo "layernorm full rc=$?"
ls -la gpurun_out/ | grep ${TAG}_
