#!/bin/bash
# One GPU session: per-kernel parity, forward/game parity, smoke, bench (both arms).  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia-smi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
echo "== gpu tests"; timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_gpu.log
echo "== smoke";     timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t_smoke.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_smoke.log
echo "== bench";     SECONDS=0; timeout 1500 python bench.py "$@" > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$? in ${SECONDS}s"; tail -1 gpurun_out/bench.log | cut -c1-400; tail -3 gpurun_out/bench.err
echo "== reference arm"; SECONDS=0; timeout 900 python bench.py --impl reference "$@" > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "rc=$? in ${SECONDS}s"; tail -1 gpurun_out/bench_ref.log | cut -c1-600
