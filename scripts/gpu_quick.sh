#!/bin/bash
# Quick GPU session: kernel parity + microbench.  Usage: bash scripts/gpu_quick.sh <tag> [microbench args]
TAG=${1:-q}; shift
mkdir -p gpurun_out
echo "== kernels"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x > gpurun_out/${TAG}_kernels.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/${TAG}_kernels.log
echo "== microbench"; timeout 900 python scripts/microbench.py "$@" > gpurun_out/${TAG}_micro.log 2>&1; echo "rc=$?"; cat gpurun_out/${TAG}_micro.log
