#!/usr/bin/env python3
"""One K1 shape, timed: python scripts/k1_probe.py N C [f16|f32|c8|x3] [P]   (for ncu: -k regex:aggregate)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapley_vit_b200 import ops  # noqa: E402

N, Cn = int(sys.argv[1]), int(sys.argv[2])
dt = torch.float32 if len(sys.argv) > 3 and sys.argv[3] == "f32" else torch.float16
P = int(sys.argv[4]) if len(sys.argv) > 4 else 21_669_514
stride = (P + 63) // 64 * 64
torch.manual_seed(N)
deltas = torch.randn(N, stride, device="cuda") * 0.02
w0 = torch.randn(stride, device="cuda") * 0.02
masks = torch.rand(Cn, N) < 0.5
masks[:, 0] |= ~masks.any(dim=1)
n = torch.arange(1, N + 1, dtype=torch.float64) * 1000
ratios = (masks * n / (masks * n).sum(dim=1, keepdim=True)).float()
kind = sys.argv[3] if len(sys.argv) > 3 else "f16"
if kind in ("c8", "x3"):
    from shapley_vit_b200 import _lib
    out = ops.OperandArray((Cn, stride), torch.float16, _lib.FMT_C8 if kind == "c8" else _lib.FMT_X3, "cuda")
    out_es = 4
else:
    out = torch.empty(Cn, stride, dtype=dt, device="cuda")
    out_es = out.element_size()
fn = lambda: ops.aggregate(deltas, w0, ratios, out=out, P=P)
fn()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(4):
    fn()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 4
by = 4.0 * P * (N + 1) + out_es * P * Cn
print(f"K1 N={N} C={Cn} {kind} P={P}: {ms*1e3:.1f} us  {by/ms/1e6:.0f} GB/s  ({by/ms/1e6/6555.5:.2f} of copy peak)")
