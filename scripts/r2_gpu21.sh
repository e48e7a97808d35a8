#!/bin/bash
run() { python bench.py --val 2048 --steps 2 --warmup 2 --no-cpu-baseline --no-parity --no-e2e --no-throughput-mode "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value'], 3), {k: round(v, 1) for k, v in d['breakdown'].items() if v}, d['clocks']['sm_mhz'])"; }
timeout 600 python -m pytest tests/test_gpu_forward.py -q -m gpu -x 2>&1 | tail -2
for i in 1 2; do
  echo "== standalone LN"; SVIT_GEMM_NO_LN_TAIL=1 run
  echo "== LN tail (staggered n)"; run
done
for s in proj down; do
echo "== probe $s (reduce GEMM without tail, for reference)"; timeout 300 python scripts/gemm_probe2.py f16c8 8 128 40 $s 2>&1 | grep "gemm2.*$s"
done
