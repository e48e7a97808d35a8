#!/bin/bash
# split attention without the trace instrumentation: one thread per row (in-tree) vs two threads per row (ab/libsvit_16warp.so)
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention_split" 2>&1 | tail -2
SVIT_LIB=$PWD/ab/libsvit_16warp.so timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py -q -m gpu -x -k "attention_split or cfg1 or geometry" 2>&1 | tail -2
for i in 1 2 3; do
  echo "== 8 warps";  timeout 300 python scripts/attn_split_probe.py 1024 197 c8 40 | tail -1
  echo "== 16 warps"; SVIT_LIB=$PWD/ab/libsvit_16warp.so timeout 300 python scripts/attn_split_probe.py 1024 197 c8 40 | tail -1
done
run() { python bench.py --val 2048 --steps 2 --warmup 2 --no-cpu-baseline --no-parity --no-e2e --no-throughput-mode "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value'], 3), {k: round(v, 1) for k, v in d['breakdown'].items() if v}, d['clocks']['sm_mhz'])"; }
for i in 1 2; do
  echo "== bench 8 warps"; run
  echo "== bench 16 warps"; SVIT_LIB=$PWD/ab/libsvit_16warp.so run
done
