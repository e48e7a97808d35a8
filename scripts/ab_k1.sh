#!/bin/bash
# A/B of K1 between ab/libsvit_old.so and the in-tree build, standalone and inside the bench step (one GPU session).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "aggregate" 2>&1 | tail -3
for i in 1 2; do
  echo "== old ($i)"; SVIT_LIB=$PWD/ab/libsvit_old.so timeout 600 python scripts/microbench.py agg 2>&1 | grep "^agg"
  echo "== new ($i)"; timeout 600 python scripts/microbench.py agg 2>&1 | grep "^agg"
done
for tag in old new; do
  if [ $tag = old ]; then export SVIT_LIB=$PWD/ab/libsvit_old.so; else unset SVIT_LIB; fi
  timeout 600 python bench.py --val 2048 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('$tag in-bench K1', d['roofline_aggregate']['achieved'], d['roofline_aggregate']['frac'], 'value', d['value'], d['clocks']['sm_mhz'])"
done
