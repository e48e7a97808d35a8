#!/bin/bash
# One GPU session: per-kernel parity, forward/game parity, smoke, bench.  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
echo "== safe kernels"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not tcgen05" > gpurun_out/t1_kernels_safe.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t1_kernels_safe.log
echo "== tcgen05";      timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "tcgen05" > gpurun_out/t2_tcgen05.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t2_tcgen05.log
echo "== forward";      timeout 1200 python -m pytest tests/test_gpu_forward.py tests/test_gpu_properties.py -q -m gpu -s > gpurun_out/t3_forward.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/t3_forward.log
echo "== smoke";        timeout 300 python __graft_entry__.py --smoke > gpurun_out/t4_smoke.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t4_smoke.log
echo "== bench";        timeout 1200 python bench.py --steps 2 --warmup 1 "$@" > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
