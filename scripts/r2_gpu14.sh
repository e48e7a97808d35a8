#!/bin/bash
# 8 ranks over NCCL: the default line (weak scaling value + e2e) with time_to_shapley_s (strong scaling, whole 255-coalition job)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 2 --warmup 2 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
echo "rc=$?"; tail -3 gpurun_out/r2_bench_8gpu.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_8gpu.json'):
    if l.startswith('{'):
        d = json.loads(l)
        print(d['value'], d['e2e']['value'], d['clocks'], json.dumps(d['time_to_shapley'])[:700])
PY
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
