#!/usr/bin/env python3
"""What a WRITE-dominated stream reaches on this GPU, next to the copy peak the K1 roofline uses: fill (write only),
copy (read + write, the MEASURED_PEAKS figure), and a 1-read / 8-write fan-out (K1's mix at C = 8 coalitions ... 128)."""
import torch

n = 1 << 30   # 1 Gi fp16 elements = 2 GiB
a = torch.empty(n, dtype=torch.float16, device="cuda")
b = torch.empty(n, dtype=torch.float16, device="cuda")


def timeit(fn, iters=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


ms = timeit(lambda: a.fill_(1.0))
print(f"fill  (write only)     : {2 * n / ms / 1e6:7.0f} GB/s")
ms = timeit(lambda: b.copy_(a))
print(f"copy  (read + write)   : {4 * n / ms / 1e6:7.0f} GB/s")
ms = timeit(lambda: a.sum())
print(f"sum   (read only)      : {2 * n / ms / 1e6:7.0f} GB/s")
src = torch.empty(n // 8, dtype=torch.float16, device="cuda")
dst = b.view(8, n // 8)
ms = timeit(lambda: dst.copy_(src.unsqueeze(0).expand(8, -1)))
print(f"fan-out 1 read 8 writes: {(2 * n // 8 + 2 * n) / ms / 1e6:7.0f} GB/s")
