#!/bin/bash
# N1: K-extension GEMM + shared-weight LoRA forward
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -s -k "k_extension" 2>&1 | tail -20
timeout 900 python -m pytest tests/test_lora.py tests/test_entrypoint.py -q -m gpu -x -s 2>&1 | tail -25
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_forward.py -q -m gpu -x -k "gemm or forward or cfg1" 2>&1 | tail -3
