#!/bin/bash
# GEMM A-operand L2 prefetch: correctness tests, then the per-shape probe and the short bench for several look-ahead depths
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py tests/test_entrypoint.py -q -m gpu -x -k "gemm or forward or cfg1 or geometry or checkpoints" 2>&1 | tail -3
for pf in 0 8 0 4 16 32; do
  echo "== prefetch $pf"
  SVIT_GEMM_PREFETCH=$pf timeout 600 python scripts/gemm_probe2.py f16c8 8 128 40 2>&1 | grep gemm2
done
for pf in 0 8 16; do
  echo "== bench prefetch $pf"
  SVIT_GEMM_PREFETCH=$pf timeout 900 python bench.py --val 2048 --steps 2 --warmup 2 --no-cpu-baseline --no-parity --no-e2e --no-throughput-mode 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value'], 3), d['breakdown'], d['clocks']['sm_mhz'])"
done
SVIT_GEMM_PREFETCH=8 timeout 600 python scripts/gemm_probe2.py f16 8 128 40 2>&1 | grep gemm2
SVIT_GEMM_PREFETCH=0 timeout 600 python scripts/gemm_probe2.py f16 8 128 40 2>&1 | grep gemm2
