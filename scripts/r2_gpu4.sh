#!/bin/bash
# REDUCE-mode GEMM with 5 stages / single staging tile vs 4 stages / double staging; image-chunk sizes (tile-wave tails)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "gemm" 2>&1 | tail -2
SVIT_GEMM_REDUCE_STAGES=5 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py -q -m gpu -x -k "gemm or cfg1 or geometry" 2>&1 | tail -2
for rs in 4 5 4 5; do
  echo "== reduce stages $rs"
  SVIT_GEMM_REDUCE_STAGES=$rs timeout 600 python scripts/gemm_probe2.py f16c8 8 128 40 proj,down 2>&1 | grep gemm2
done
SVIT_GEMM_REDUCE_STAGES=5 timeout 600 python scripts/gemm_probe2.py f16 8 128 40 proj,down 2>&1 | grep gemm2
SVIT_GEMM_REDUCE_STAGES=4 timeout 600 python scripts/gemm_probe2.py f16 8 128 40 proj,down 2>&1 | grep gemm2
run() { python bench.py --val 2000 --steps 2 --warmup 2 --no-cpu-baseline --no-parity --no-e2e --no-throughput-mode "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(round(d['value'], 3), {k: round(v, 1) for k, v in d['breakdown'].items() if v}, d['clocks']['sm_mhz'])"; }
for ch in 128 200 250 128 200; do
  echo "== bench chunk $ch (4 stages)"; run --image-chunk $ch
done
echo "== bench chunk 200 (5 stages)"; SVIT_GEMM_REDUCE_STAGES=5 run --image-chunk 200
echo "== bench chunk 128 (5 stages)"; SVIT_GEMM_REDUCE_STAGES=5 run --image-chunk 128
