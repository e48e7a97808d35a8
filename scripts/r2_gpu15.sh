#!/bin/bash
# e2e with chunk-level upload readiness, smoke(), HBM write probe
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
python scripts/hbm_write_probe.py 2>&1 | tail -5
python bench.py --no-cpu-baseline --no-parity --no-throughput-mode > gpurun_out/r2_e2e_check.json 2> gpurun_out/r2_e2e_check.err; echo "rc=$?"; tail -2 gpurun_out/r2_e2e_check.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_e2e_check.json'):
    if l.startswith('{'):
        d = json.loads(l); print('value', d['value'], 'e2e', d['e2e']['value'], 'ratio', d['e2e']['value'] / d['value'], d['clocks']['sm_mhz'])
PY
timeout 600 python -m pytest tests/test_gpu_forward.py tests/test_entrypoint.py -q -m gpu -x 2>&1 | tail -2
