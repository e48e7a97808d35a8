#!/bin/bash
# 2 ranks over NCCL: the packed on-device all-gather, shared seeds, time_to_shapley_s with the cross-rank identity check
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --val 1024 --steps 2 --warmup 1 --time-to-shapley > gpurun_out/r2_2gpu_val1024.json 2> gpurun_out/r2_2gpu_val1024.err
echo "rc=$?"; tail -3 gpurun_out/r2_2gpu_val1024.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_2gpu_val1024.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print(d['value'], d['e2e']['value'], json.dumps(d['time_to_shapley'])[:900])
PY
# stochastic estimator under 2 ranks with seed None (shared seed) through the public entry point
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/dist_entry_check.py 2>&1 | tail -6
