#!/usr/bin/env python3
"""Attention kernel, a few launches: the command ncu wraps.   python scripts/attn_probe.py [n_seq] [T]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shapley_vit_b200 import ops
n_seq = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 197
qkv = (torch.randn(n_seq, T, 2304, device="cuda") * 0.5).half()
for _ in range(3):
    ops.attention(qkv, 12)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    ops.attention(qkv, 12)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(f"attention n_seq={n_seq} T={T}: {ms*1e3:.1f} us {4.0*n_seq*T*T*768/ms/1e9:.1f} TFLOP/s")
