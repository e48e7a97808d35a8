#!/bin/bash
# 4 ranks: the default line with time_to_shapley_s (strong scaling) for the scaling table
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 2 --warmup 2 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err
echo "rc=$?"; tail -2 gpurun_out/r2_bench_4gpu.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_4gpu.json'):
    if l.startswith('{'):
        d = json.loads(l)
        t = d['time_to_shapley']
        print(d['value'], d['e2e']['value'], d['clocks']['sm_mhz'], t['value'], t['ideal_s'], t['vs_ideal'], t['cross_rank_identity']['bit_identical'])
PY
