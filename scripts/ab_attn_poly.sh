#!/bin/bash
# share of exp2 evaluated on the FMA pipes (polynomial) instead of MUFU in the tcgen05 attention: 1/2 (old), 1/3, 1/4
for i in 1 2; do
for v in old p3 p4; do echo -n "$v: "; SVIT_LIB=$PWD/ab/libsvit_$v.so timeout 300 python scripts/microbench.py attn 2>&1 | grep "float16"; done
done
SVIT_LIB=$PWD/ab/libsvit_p4.so timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "attention" 2>&1 | tail -2
