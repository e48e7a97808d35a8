#!/bin/bash
# BASELINE configs 3 and 4 with the round-2 kernels (bounded budgets, real geometry)
mkdir -p gpurun_out
timeout 900 python scripts/cfg3_run.py --perms 200 --val 500 > gpurun_out/r2_cfg3_mc.json 2> gpurun_out/r2_cfg3_mc.err; echo "cfg3 rc=$?"; tail -c 600 gpurun_out/r2_cfg3_mc.json
for prec in bf16 f16c8; do
  timeout 900 python bench.py --vit large --clients 10 --coalition-batch 32 --image-chunk 32 --val 1000 --precision $prec --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-throughput-mode > gpurun_out/r2_bench_cfg4_vitl_$prec.json 2> gpurun_out/r2_bench_cfg4_$prec.err; echo "cfg4 $prec rc=$?"; tail -2 gpurun_out/r2_bench_cfg4_$prec.err
  python - <<PY
import json
for l in open("gpurun_out/r2_bench_cfg4_vitl_$prec.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$prec", d["value"], d["e2e"]["value"], d["roofline"].get("tensor_pipe_frac", d["roofline"]["frac"]), d["roofline_aggregate"]["frac"], d["clocks"]["sm_mhz"])
PY
done
