#!/bin/bash
# f16x3 (split-precision tensor-core mode): parity tests, microbench, then a bench line at reduced validation size.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -s -k "f16x3 or split or attention" 2>&1 | grep -v "^$" | tail -25
timeout 1200 python -m pytest tests/test_gpu_forward.py -q -m gpu -x -s -k "f16x3 or f32" 2>&1 | grep -v "^$" | tail -15
timeout 300 python scripts/microbench.py attn32 2>&1 | tail -3
timeout 1200 python bench.py --precision f16x3 --val ${VAL:-1024} --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/f16x3_bench.json 2> gpurun_out/f16x3_bench.err; echo "rc=$?"; tail -3 gpurun_out/f16x3_bench.err; cat gpurun_out/f16x3_bench.json
