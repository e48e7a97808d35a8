#!/bin/bash
# Split attention with two threads per row (16 softmax warps): tests, probe, trace, short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py -q -m gpu -x -k "attention_split or forward or cfg1 or geometry" 2>&1 | tail -3
timeout 300 python scripts/attn_split_probe.py 1024 197 c8 40 | tail -1
timeout 300 python scripts/attn_split_probe.py 1024 197 x3 40 | tail -1
timeout 300 python scripts/attn_split_trace.py | head -14
run() { python bench.py --val 2048 --steps 2 --warmup 2 --no-cpu-baseline --no-parity --no-e2e --no-throughput-mode "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value'], 3), {k: round(v, 1) for k, v in d['breakdown'].items() if v}, d['clocks']['sm_mhz'])"; }
run; run
