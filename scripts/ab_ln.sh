#!/bin/bash
# LayerNorm: fixed-size persistent kernel vs the generic one (SVIT_LN_GENERIC=1), standalone and inside the bench step
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "layernorm" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_forward.py -q -m gpu -x 2>&1 | tail -2
for g in 1 0 1 0; do echo -n "SVIT_LN_GENERIC=$g "; SVIT_LN_GENERIC=$g python scripts/microbench.py ln 2>&1 | grep "out=float16"; done
for g in 1 0; do
  SVIT_LN_GENERIC=$g timeout 600 python bench.py --val 2048 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); b=d['breakdown']; print('generic=$g in-bench LN GB/s', b['layernorm_gbs'], 'ln_ms', b['layernorm_ms'], 'value', d['value'], d['clocks']['sm_mhz'])"
done
