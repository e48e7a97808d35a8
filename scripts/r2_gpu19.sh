#!/bin/bash
# full-size validation sets (10 000 images) for configs 3 and 4, bounded in the number of coalitions
mkdir -p gpurun_out
timeout 1200 python scripts/cfg3_run.py --perms 16 --val 10000 > gpurun_out/r2_cfg3_mc_full_val.json 2> gpurun_out/r2_cfg3_full.err; echo "cfg3 rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_cfg3_mc_full_val.json')); mc = d['monte_carlo']
print({k: v for k, v in mc.items() if not isinstance(v, list)})
PY
timeout 1500 python bench.py --vit large --clients 10 --coalition-batch 32 --image-chunk 32 --val 10000 --precision bf16 --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-throughput-mode --no-e2e > gpurun_out/r2_bench_cfg4_vitl_bf16_full_val.json 2> gpurun_out/r2_cfg4_full.err; echo "cfg4 rc=$?"; tail -2 gpurun_out/r2_cfg4_full.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_cfg4_vitl_bf16_full_val.json'):
    if l.startswith('{'):
        d = json.loads(l); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline_aggregate']['frac'], d['clocks']['sm_mhz'])
PY
