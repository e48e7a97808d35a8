#!/bin/bash
# does the phase-trace instrumentation cost anything?  in-tree build (trace hooks) vs ab/libsvit_notrace.so, one session
for i in 1 2 3; do
  echo "== trace hooks"; timeout 300 python scripts/attn_split_probe.py 1024 197 c8 40 | tail -1
  echo "== no hooks";    SVIT_LIB=$PWD/ab/libsvit_notrace.so timeout 300 python scripts/attn_split_probe.py 1024 197 c8 40 | tail -1
done
