#!/usr/bin/env python3
"""BASELINE config 5: aggregation-only bandwidth sweep.  4-64 clients x ViT-S/B/L parameter stacks x coalition
batch 1-128, fp16, C8 (the default precision's operand feed) and fp32 (exact) outputs; achieved GB/s = algorithmic bytes / CUDA-event time,
against the measured copy peak (MEASURED_PEAKS.json).   python scripts/k1_sweep.py > profiles/r2_k1_sweep.md"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapley_vit_b200 import ops  # noqa: E402

peak = 6555.5
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
SIZES = {"ViT-S": 21_669_514, "ViT-B": 85_806_346, "ViT-L": 303_311_882}


def timeit(fn, iters=4, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


print("# K1 aggregation sweep (BASELINE config 5)\n")
print(f"`{torch.cuda.get_device_name(0)}`; algorithmic bytes = 4 P (N+1) + s_out P C; fraction of the measured copy peak "
      f"{peak:.0f} GB/s; random 50 % membership, FedAvg ratios; CUDA events, 4 launches after 1 warm-up.\n")
for model, P in SIZES.items():
    stride = (P + 63) // 64 * 64
    print(f"## {model}: P = {P:,}\n")
    print("| N | out | " + " | ".join(f"C={c}" for c in (1, 2, 4, 8, 16, 32, 64, 128)) + " |")
    print("|---|---|" + "---:|" * 8)
    for N in (4, 8, 16, 32, 64):
        if (N + 1) * stride * 4 > 60e9:
            continue
        torch.manual_seed(N)
        deltas = torch.randn(N, stride, device="cuda") * 0.02
        w0 = torch.randn(stride, device="cuda") * 0.02
        for kind in ("f16", "c8", "f32"):
            dt = torch.float32 if kind == "f32" else torch.float16
            es = 2 if kind == "f16" else 4
            cells = []
            for Cn in (1, 2, 4, 8, 16, 32, 64, 128):
                if Cn * stride * es + (N + 1) * stride * 4 > 150e9 or (kind == "f32" and Cn > 32):
                    cells.append("—")
                    continue
                masks = torch.rand(Cn, N) < 0.5
                masks[:, 0] |= ~masks.any(dim=1)
                n = torch.arange(1, N + 1, dtype=torch.float64) * 1000
                ratios = (masks * n / (masks * n).sum(dim=1, keepdim=True)).float()
                if kind == "c8":   # the default precision's operand feed: fp16 + two e4m3 planes, 4 bytes per element
                    from shapley_vit_b200 import _lib
                    out = ops.OperandArray((Cn, stride), torch.float16, _lib.FMT_C8, "cuda")
                else:
                    out = torch.empty(Cn, stride, dtype=dt, device="cuda")
                ms = timeit(lambda: ops.aggregate(deltas, w0, ratios, out=out, P=P))
                by = 4.0 * P * (N + 1) + es * P * Cn
                gbs = by / ms / 1e6
                cells.append(f"{gbs:.0f} ({gbs / peak:.2f})")
                del out
            print(f"| {N} | {kind} | " + " | ".join(cells) + " |", flush=True)
        del deltas, w0
    print()
