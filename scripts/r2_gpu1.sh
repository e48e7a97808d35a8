#!/bin/bash
# Round-2 GPU session 1: the operand-plane refactor (f16x3 planes, f16c8 compensation, tcgen05 split attention).
# Separate pytest processes per group: a trapped kernel poisons its CUDA context only.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
run() { tag=$1; shift; echo "== $tag"; timeout 900 "$@" > gpurun_out/g1_$tag.log 2>&1; echo "rc=$?"; tail -${TAILN:-6} gpurun_out/g1_$tag.log; }
TAILN=25 run gemm_split python -m pytest tests/test_gpu_kernels.py -q -m gpu -s -k "f16x3 or f16c8 or split_plane or split_operand" 
TAILN=25 run attn_split python -m pytest tests/test_gpu_kernels.py -q -m gpu -s -k "attention_split"
TAILN=12 run k1_ln python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "aggregate or layernorm or patchify"
TAILN=12 run kernels_rest python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not (f16x3 or f16c8 or split_plane or split_operand or attention_split or aggregate or layernorm or patchify)"
TAILN=30 run forward python -m pytest tests/test_gpu_forward.py -q -m gpu -s
for prec in f16c8 f16x3 f16; do
  TAILN=1 run bench_$prec python bench.py --precision $prec --val 2048 --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-e2e
done
