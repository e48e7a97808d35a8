#!/bin/bash
# K1 fused fold for C8, cheaper split_c8 in every producer, coalesced [CLS] split attention: tests, probes, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py tests/test_lazy_rounds.py tests/test_gpu_configs.py tests/test_lora.py -q -m gpu -x 2>&1 | tail -4
for a in "8 8 c8 85806346" "8 8 f16 85806346" "8 8 x3 85806346" "8 8 f32 85806346" "16 16 c8 85806346" "8 32 c8 85806346"; do timeout 300 python scripts/k1_probe.py $a | tail -1; done
timeout 300 python scripts/attn_split_probe.py 1024 197 c8 40 | tail -1
run() { python bench.py --val 2048 --steps 2 --warmup 2 --no-cpu-baseline --no-parity --no-e2e --no-throughput-mode "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value'], 3), {k: round(v, 1) for k, v in d['breakdown'].items() if v}, d['clocks']['sm_mhz'], 'K1', round(d['roofline_aggregate']['frac'], 3))"; }
run; run
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-throughput-mode --val 1024 > gpurun_out/r2b_ncu_launches.log 2>&1
echo "launch list rc=$?"
