#!/bin/bash
# final kernels of round 2: the whole GPU suite, the default bench line, the reference arm, launch list + full captures, K1 sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
python bench.py > gpurun_out/r2f_bench_default.json 2> gpurun_out/r2f_bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2f_bench_default.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err; echo "reference rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-throughput-mode --val 1024"
$CMD > gpurun_out/r2f_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2f_plain.log; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_tc2_kernel -s 29 -c 4 -o gpurun_out/r2f_gemm -f $CMD > gpurun_out/r2f_ncu_gemm.log 2>&1; echo "gemm full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:aggregate -c 2 -o gpurun_out/r2f_aggregate -f $CMD > gpurun_out/r2f_ncu_agg.log 2>&1; echo "aggregate full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:attention_tc_split -s 4 -c 1 -o gpurun_out/r2f_attention -f $CMD > gpurun_out/r2f_ncu_att.log 2>&1; echo "attention full rc=$?"
timeout 900 python scripts/k1_sweep.py > gpurun_out/r2_k1_sweep.md 2> gpurun_out/r2_k1_sweep.err; echo "sweep rc=$?"
