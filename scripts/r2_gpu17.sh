#!/bin/bash
# 4-CTA cluster GEMM (two pairs share the B tile by TMA multicast): correctness, then the per-shape probe against the pair kernel
mkdir -p gpurun_out
SVIT_GEMM_CLUSTER4=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "gemm" 2>&1 | tail -5
SVIT_GEMM_CLUSTER4=1 timeout 600 python -m pytest tests/test_gpu_forward.py -q -m gpu -x 2>&1 | tail -3
for i in 1 2; do
  echo "== pair"; timeout 600 python scripts/gemm_probe2.py f16c8 8 128 40 2>&1 | grep gemm2
  echo "== quad"; SVIT_GEMM_CLUSTER4=1 timeout 600 python scripts/gemm_probe2.py f16c8 8 128 40 2>&1 | grep gemm2
done
echo "== pair f16"; timeout 600 python scripts/gemm_probe2.py f16 8 128 40 2>&1 | grep gemm2
echo "== quad f16"; SVIT_GEMM_CLUSTER4=1 timeout 600 python scripts/gemm_probe2.py f16 8 128 40 2>&1 | grep gemm2
