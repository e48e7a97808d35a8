#!/bin/bash
# GEMM microbench, ab/libsvit_old.so vs the in-tree build, one GPU session (+ the parity tests that cover the epilogues)
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py -q -m gpu -x -k "gemm or forward or cfg1 or geometry" 2>&1 | tail -3
for i in 1 2; do
  echo "== old ($i)"; SVIT_LIB=$PWD/ab/libsvit_old.so timeout 600 python scripts/microbench.py gemm 2>&1 | grep "gemm\[f16\]"
  echo "== new ($i)"; timeout 600 python scripts/microbench.py gemm 2>&1 | grep "gemm\[f16\]"
done
