#!/usr/bin/env python3
"""One GEMM shape, a few launches: the command ncu wraps.   python scripts/gemm_probe.py qkv|proj|up|down|plain [C] [B]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapley_vit_b200 import _lib, ops

which = sys.argv[1] if len(sys.argv) > 1 else "plain"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 8
B = int(sys.argv[3]) if len(sys.argv) > 3 else 128
P = _lib.PRECISIONS["f16"]; dt = torch.float16; T, h, ff = 197, 768, 3072; M = B * T
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *s, dtype=dt, scale=0.05: (torch.randn(*s, device="cuda", generator=g) * scale).to(dtype)
X = rnd(C, M, h, dtype=torch.float32, scale=1.0)
cfg = {
    "plain": (h, 3 * h, dict()),
    "qkv": (h, 3 * h, dict(bias=rnd(C, 3 * h, dtype=torch.float32))),
    "proj": (h, h, dict(bias=rnd(C, h, dtype=torch.float32), residual=X, out=X, out_dtype=torch.float32)),
    "up": (h, ff, dict(bias=rnd(C, ff, dtype=torch.float32), gelu=True)),
    "down": (ff, h, dict(bias=rnd(C, h, dtype=torch.float32), residual=X, out=X, out_dtype=torch.float32)),
}[which]
K, N, kw = cfg
A, W = rnd(C, M, K), rnd(C, N, K)
if "out" not in kw:
    kw["out"] = torch.empty(C, M, N, dtype=dt, device="cuda")
for _ in range(4):
    ops.gemm(P, A, W, **kw)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(4):
    ops.gemm(P, A, W, **kw)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 4
print(f"{which}: {ms*1e3:.1f} us {2.0*C*M*N*K/ms/1e9:.1f} TFLOP/s")
