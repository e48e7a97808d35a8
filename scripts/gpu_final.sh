#!/bin/bash
# Round-end verification: the whole GPU suite, smoke, both bench arms, then the config-4 geometry line.
bash scripts/gpu_round.sh
echo "== cfg4 geometry (ViT-L/16 bf16, 10 clients, 32 models per GEMM group)"
SECONDS=0; timeout 1200 python bench.py --vit large --clients 10 --coalition-batch 32 --precision bf16 --val 1000 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cfg4.log 2> gpurun_out/bench_cfg4.err; echo "rc=$? in ${SECONDS}s"; tail -1 gpurun_out/bench_cfg4.log | cut -c1-600; tail -3 gpurun_out/bench_cfg4.err
echo "== f16x3 (split-precision mode), full size"
SECONDS=0; timeout 1200 python bench.py --precision f16x3 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_f16x3.log 2> gpurun_out/bench_f16x3.err; echo "rc=$? in ${SECONDS}s"; tail -1 gpurun_out/bench_f16x3.log | cut -c1-300; tail -3 gpurun_out/bench_f16x3.err
