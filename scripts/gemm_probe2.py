#!/usr/bin/env python3
"""The four GEMM shapes of one encoder layer on pre-split operands, per precision, timed with CUDA events.
    python scripts/gemm_probe2.py [precision=f16c8] [C=8] [B=128] [reps=20] [which=qkv,proj,up,down]
Each shape: warm-up, then `reps` back-to-back launches (long enough to sit at the sustained clock)."""
import ctypes as C_
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapley_vit_b200 import _lib, ops
from shapley_vit_b200._lib import EpilogueC, check

prec_name = sys.argv[1] if len(sys.argv) > 1 else "f16c8"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 8
B = int(sys.argv[3]) if len(sys.argv) > 3 else 128
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
which = (sys.argv[5] if len(sys.argv) > 5 else "qkv,proj,up,down").split(",")
P = _lib.PRECISIONS[prec_name]
fmt = _lib.OPERAND_FORMAT[P]
odt = ops.TORCH_DTYPE[_lib.OPERAND_DTYPE[P]]
T, h, ff = 197, 768, 3072
M = B * T
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s, scale=0.05: torch.randn(*s, device="cuda", generator=g) * scale
lib = _lib.load()
X = rnd(C, M, h, scale=1.0)
shapes = {"qkv": (h, 3 * h, False, False), "proj": (h, h, True, False), "up": (h, ff, False, True), "down": (ff, h, True, False)}
tot_ms, tot_fl = 0.0, 0.0
for name in which:
    K, N, resid, gelu = shapes[name]
    A = ops.OperandArray.from_float(rnd(C, M, K), fmt, odt)
    W = ops.OperandArray.from_float(rnd(C, N, K), fmt, odt)
    bias = rnd(C, N)
    epi = EpilogueC()
    epi.bias, epi.bias_gs = bias.data_ptr(), N
    epi.gelu = int(gelu)
    if resid:
        out_ptr, out_dt = X.data_ptr(), _lib.F32
        epi.residual, epi.residual_gs = X.data_ptr(), M * N
        keep = X
    else:
        keep = ops.OperandArray((C, M, N), odt, fmt, "cuda")
        out_ptr, out_dt = keep.ptr, _lib.OPERAND_DTYPE[P]
    st = C_.c_void_p(torch.cuda.current_stream().cuda_stream)
    call = lambda: check(lib.svit_gemm(P, C_.c_void_p(A.ptr), M * K, C_.c_void_p(W.ptr), N * K, C_.c_void_p(out_ptr), M * N, out_dt,
                                       C, M, N, K, C_.byref(epi), st))
    for _ in range(max(3, reps // 2)):
        call()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        call()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    fl = 2.0 * C * M * N * K
    tot_ms += ms
    tot_fl += fl
    print(f"gemm2[{prec_name}] {name:5s}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s (algorithmic)")
    del A, W, keep
print(f"gemm2[{prec_name}] layer: {tot_ms * 1e3:8.1f} us  {tot_fl / tot_ms / 1e9:7.1f} TFLOP/s (algorithmic)")
