#!/bin/bash
# the bench step with ab/libsvit_old.so and with the in-tree build, alternating, one GPU session
for i in 1 2; do
for tag in old new; do
  if [ $tag = old ]; then export SVIT_LIB=$PWD/ab/libsvit_old.so; else unset SVIT_LIB; fi
  timeout 600 python bench.py --val ${VAL:-2048} --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); b=d['breakdown']; print('$tag value', round(d['value'],3), 'gemm_ms', round(b['gemm_ms'],1), 'gemm TF/s', round(d['roofline']['achieved']), 'clock', d['clocks']['sm_mhz'], 'W', d['clocks']['power_w_max'])"
done; done
