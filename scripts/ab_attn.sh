#!/bin/bash
# A/B of two builds of libsvit in ONE gpurun call (box-to-box variation is +-5 %): ab/libsvit_old.so vs the in-tree build.
# Usage: bash scripts/ab_attn.sh <microbench args...>
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention" 2>&1 | tail -3
for i in 1 2 3; do
  echo "== old ($i)"; SVIT_LIB=$PWD/ab/libsvit_old.so timeout 600 python scripts/microbench.py "$@" 2>&1 | tail -6
  echo "== new ($i)"; timeout 600 python scripts/microbench.py "$@" 2>&1 | tail -6
done
