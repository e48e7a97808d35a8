#!/bin/bash
# LayerNorm tail in the residual GEMMs: parity suite, then A/B against the standalone LayerNorm (SVIT_GEMM_NO_LN_TAIL=1)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
run() { python bench.py --val 2048 --steps 2 --warmup 2 --no-cpu-baseline --no-parity --no-e2e --no-throughput-mode "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value'], 3), {k: round(v, 1) for k, v in d['breakdown'].items() if v}, d['clocks']['sm_mhz'], d['gpu_launches'])"; }
for i in 1 2; do
  echo "== standalone LN"; SVIT_GEMM_NO_LN_TAIL=1 run
  echo "== LN tail"; run
done
echo "== f16 standalone"; SVIT_GEMM_NO_LN_TAIL=1 run --precision f16
echo "== f16 LN tail"; run --precision f16
