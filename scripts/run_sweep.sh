for cfg in "8 128" "8 256" "16 128" "4 256" "8 64"; do set -- $cfg; python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity --val 2048 --coalition-batch $1 --image-chunk $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('Cb=$1 B=$2', round(d['value'],3), 'gemm TF/s', round(d['roofline']['achieved'],1), d['breakdown']['forward_ms'], d['clocks']['sm_mhz'])"; done
