#!/usr/bin/env python3
"""Split-precision attention kernel, back-to-back launches.   python scripts/attn_split_probe.py [n_seq] [T] [c8|x3] [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapley_vit_b200 import _lib, ops
n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 197
fmt = _lib.FMT_C8 if (len(sys.argv) <= 3 or sys.argv[3] == "c8") else _lib.FMT_X3
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
qkv = torch.randn(n_seq, T, 2304, device="cuda") * 0.5
for _ in range(3):
    ops.attention_split(qkv, 12, fmt)
torch.cuda.synchronize()
# time the kernel alone: pre-split the input once, then call the library directly
import ctypes as C
from shapley_vit_b200._lib import check
q3 = ops.OperandArray.from_float(qkv, _lib.FMT_X3)
out = ops.OperandArray((n_seq, T, 768), torch.float16, fmt, "cuda")
lib = _lib.load()
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
call = lambda: check(lib.svit_attention_split(C.c_void_p(q3.ptr), C.c_void_p(out.ptr), fmt, n_seq, T, 12, 64, st))
for _ in range(5):
    call()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    call()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(f"attention_split n_seq={n_seq} T={T}: {ms*1e3:.1f} us {4.0*n_seq*T*T*768/ms/1e9:.1f} TFLOP/s (algorithmic)")
