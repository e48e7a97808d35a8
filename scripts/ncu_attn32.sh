#!/bin/bash
# fp32 attention: plain timing, then one ncu --set full capture.  Usage: bash scripts/ncu_attn32.sh <tag>
TAG=${1:-a32}; mkdir -p gpurun_out
python scripts/microbench.py attn32 > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:attention_rt -s 2 -c 1 -o gpurun_out/${TAG}_attn32 -f python scripts/microbench.py attn32 > gpurun_out/${TAG}_ncu.log 2>&1
cat gpurun_out/${TAG}_plain.log
