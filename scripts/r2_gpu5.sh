#!/bin/bash
# Split attention: MMA issue order S0,S1,O0,O1 (new) against the interleaved order, one session
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py -q -m gpu -x -k "attention_split or forward or cfg1 or geometry" 2>&1 | tail -2
for i in 1 2; do
  echo "== interleaved"; SVIT_ATTN_SPLIT_INTERLEAVED=1 timeout 300 python scripts/attn_split_probe.py 1024 197 c8 40 | tail -1
  echo "== S0 S1 O0 O1"; timeout 300 python scripts/attn_split_probe.py 1024 197 c8 40 | tail -1
done
run() { python bench.py --val 2048 --steps 2 --warmup 2 --no-cpu-baseline --no-parity --no-e2e --no-throughput-mode "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value'], 3), {k: round(v, 1) for k, v in d['breakdown'].items() if v}, d['clocks']['sm_mhz'])"; }
for i in 1 2; do
  echo "== bench interleaved"; SVIT_ATTN_SPLIT_INTERLEAVED=1 run
  echo "== bench new order"; run
done
